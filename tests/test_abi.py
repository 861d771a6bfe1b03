"""CPU: the C-ABI library loads and exports every symbol include/alignq_b200.h declares, the ctypes
binding lists exactly those symbols, and the host-only helpers work.  No compute calls."""
import ctypes
import os
import re

import pytest

import alignq_b200 as aq
from alignq_b200 import _lib as L

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(REPO, "include", "alignq_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(alignq_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_in_tree():
    assert os.path.exists(L.LIB_PATH), "run __graft_entry__.build() first"
    assert os.path.dirname(L.LIB_PATH).startswith(REPO)


def test_exports_every_declared_symbol():
    lib = ctypes.CDLL(L.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/alignq_b200.h but not exported"


def test_binding_matches_header():
    assert sorted(L.SIGNATURES) == declared_symbols()


def test_abi_version_and_error_strings():
    lib = aq.load_library()
    assert lib.alignq_abi_version() == L.ABI_VERSION
    assert b"invalid argument" in lib.alignq_error_string(-1)
    assert lib.alignq_error_string(0) == b"ok"


def test_sgd_struct_layout_matches_header():
    # 5 pointers + int64 + 4 floats + 2 int32 = 72 bytes, 8-byte aligned
    assert ctypes.sizeof(L.SgdTensor) == 72
    assert L.SgdTensor.numel.offset == 40 and L.SgdTensor.lr.offset == 48 and L.SgdTensor.first_step.offset == 68


def test_chunk_plan_host_helper():
    seg_off, chunk_seg, seg_chunk0, nchunks = L.plan_chunks([432, 4096, 4097, 0, 9000])
    assert seg_off == [0, 432, 4528, 8625, 8625, 17625]
    assert seg_chunk0 == [0, 1, 2, 4, 4, 7] and nchunks == 7
    assert chunk_seg == [0, 1, 2, 2, 4, 4, 4]


def test_ws_bytes_is_monotone_and_capped():
    lib = aq.load_library()
    a = lib.alignq_gram_ws_bytes(128, 4096)
    b = lib.alignq_gram_ws_bytes(128, 16384)
    c = lib.alignq_gram_ws_bytes(128, 1 << 24)
    assert 0 < a <= b <= c <= (256 << 20)
    assert lib.alignq_gram_ws_bytes(0, 10) == 0


def test_header_is_plain_c_and_argument_counts_match_the_binding(tmp_path):
    """include/alignq_b200.h is the drop-in boundary: it must compile as C99 (no C++ / torch types) and every prototype
    must take exactly as many arguments as the ctypes signature table passes."""
    import re
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "t.c"
    src.write_text('#include "alignq_b200.h"\nint main(void) { return alignq_abi_version() == 0; }\n')
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(REPO, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = open(os.path.join(REPO, "include", "alignq_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = dict((m.group(1), m.group(2)) for m in re.finditer(r"\b(alignq_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S))
    assert len(protos) >= 40
    for name, (_, argtypes) in L.SIGNATURES.items():
        assert name in protos, name
        params = protos[name].strip()
        n = 0 if params in ("", "void") else len([p for p in params.split(",") if p.strip()])
        assert n == len(argtypes), f"{name}: header has {n} parameters, the binding passes {len(argtypes)}"
