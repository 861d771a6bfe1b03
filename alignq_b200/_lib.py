"""ctypes binding of ``csrc/libalignq_b200.so`` (the C ABI declared in ``include/alignq_b200.h``).

There is no CPU fallback and no alternative backend: if the library is missing or a launch
fails, the call raises.  PyTorch is used only for device memory and streams -- every pointer
handed to the library is ``tensor.data_ptr()`` of a contiguous fp32 CUDA tensor and every launch
goes to ``torch.cuda.current_stream()``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libalignq_b200.so")

ABI_VERSION = 1
CHUNK = 4096
VARIANT_ID = {"A": 0, "B": 1, "C": 2}
GRAM_MODE_ID = {"fp32": 0, "tf32x3": 1, "bf16": 2}
CONV_MODE_ID = {"tf32": 0, "tf32x3": 1}


class SgdTensor(C.Structure):
    """Mirror of ``alignq_sgd_tensor_t``."""
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("buf", C.c_void_p), ("w_cdf", C.c_void_p),
                ("w_pdf", C.c_void_p), ("numel", C.c_int64), ("lr", C.c_float), ("momentum", C.c_float),
                ("dampening", C.c_float), ("weight_decay", C.c_float), ("nesterov", C.c_int32),
                ("first_step", C.c_int32)]


_P, _I, _L, _F, _Z = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
# name -> (restype, argtypes); must list every symbol declared in include/alignq_b200.h
SIGNATURES = {
    "alignq_abi_version": (_I, []),
    "alignq_error_string": (C.c_char_p, [_I]),
    "alignq_launch_count": (C.c_uint64, []),
    "alignq_act_fwd": (_I, [_P, _P, _P, _L, _I, _F, _I, _I, _P]),
    "alignq_act_bwd": (_I, [_P, _P, _P, _L, _I, _F, _I, _I, _P]),
    "alignq_act_grad_scale": (_F, [_I, _F, _I, _I]),
    "alignq_uniform_q_fwd": (_I, [_P, _P, _L, _I, _P]),
    "alignq_cdf_fwd": (_I, [_P, _P, _P, _I, _I, _F, _P, _P, _L, _P]),
    "alignq_cdf_bwd": (_I, [_P, _P, _P, _I, _I, _F, _P, _P, _P, _L, _P]),
    "alignq_wq_plan": (_L, [_P, _I, _P, _P]),
    "alignq_wq_forward": (_I, [_P, _P, _P, _P, _I, _L, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "alignq_wq_backward": (_I, [_P, _P, _P, _P, _P, _P, _I, _L, _I, _P, _P, _I, _P, _P]),
    "alignq_gram_ws_bytes": (_Z, [_I, _L]),
    "alignq_corr_fwd": (_I, [_P, _P, _I, _L, _F, _P, _P, _Z, _I, _P]),
    "alignq_corr_bwd": (_I, [_P, _P, _P, _I, _L, _F, _P, _P, _P, _Z, _P]),
    "alignq_gram_sums_fwd": (_I, [_P, _I, _L, _F, _F, _P, _P, _Z, _I, _P]),
    "alignq_gram_sums_to_d": (_I, [_P, _P, _I, _I, _P, _P]),
    "alignq_act_bwd_add": (_I, [_P, _P, _P, _P, _L, _I, _F, _I, _I, _P]),
    "alignq_act_admm_fwd": (_I, [_P, _I, _L, _I, _F, _F, _P, _P, _I, _F, _F, _P, _P, _P, _P, _P, _Z, _I, _P]),
    "alignq_act_admm_bwd": (_I, [_P, _P, _P, _P, _I, _L, _I, _F, _F, _P, _P, _Z, _I, _P]),
    "alignq_admm_loss": (_I, [_P, _I, _P, _P, _I, _I, _F, _F, _P, _I, _P, _P, _P, _P, _P]),
    "alignq_admm_zu_update": (_I, [_P, _P, _P, _I, _I, _I, _F, _F, _P]),
    "alignq_gram_bf16_ws_bytes": (_Z, [_I]),
    "alignq_gram_bf16": (_I, [_P, _I, _L, _I, _P, _P, _Z, _P]),
    "alignq_bn_act_ws_doubles": (_Z, [_I]),
    "alignq_bn_act_fwd": (_I, [_P, _L, _I, _P, _P, _P, _P, _F, _F, _I, _I, _F, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "alignq_bn_act_bwd": (_I, [_P, _P, _P, _L, _I, _P, _P, _P, _P, _I, _I, _F, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "alignq_bn_act_bwd_apply": (_I, [_P, _P, _P, _L, _I, _P, _P, _P, _P, _I, _F, _I, _I, _P, _P, _P, _P]),
    "alignq_conv3x3_bwd_data_bnreduce": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _F, _I, _P, _P, _P, _P, _P]),
    "alignq_bn_act_bwd_sum": (_I, [_P, _P, _P, _P, _L, _I, _P, _P, _P, _P, _I, _I, _F, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "alignq_bn_act_sync_stats": (_I, [_P, _L, _I, _P, _P, _P, _P]),
    "alignq_bn_act_sync_apply": (_I, [_P, _L, _L, _I, _P, _P, _P, _P, _P, _F, _F, _I, _F, _I, _I, _P, _P, _P, _P, _P, _P]),
    "alignq_bn_act_sync_bwd_reduce": (_I, [_P, _P, _P, _L, _I, _P, _P, _P, _P, _I, _F, _I, _I, _P, _P, _P, _P, _P, _P]),
    "alignq_bn_act_sync_bwd_apply": (_I, [_P, _P, _P, _L, _L, _I, _P, _P, _P, _P, _I, _F, _I, _I, _P, _P, _P, _P, _P]),
    "alignq_bn_act_peer_bytes": (_Z, []),
    "alignq_bn_act_fwd_peer": (_I, [_P, _L, _L, _I, _P, _P, _P, _P, _F, _F, _I, _F, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "alignq_bn_act_bwd_peer": (_I, [_P, _P, _P, _L, _L, _I, _P, _P, _P, _P, _I, _F, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "alignq_bn_act_bwd_peer_sum": (_I, [_P, _P, _P, _P, _L, _L, _I, _P, _P, _P, _P, _I, _F, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "alignq_conv3x3_fwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "alignq_conv3x3_fwd_bnstats": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P]),
    "alignq_bn_act_apply": (_I, [_P, _L, _I, _P, _P, _P, _P, _I, _F, _I, _I, _P, _P, _P]),
    "alignq_conv3x3_bwd_data": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "alignq_conv3x3_ws_bytes": (_Z, [_I]),
    "alignq_head_ce_ws_bytes": (_Z, [_I]),
    "alignq_head_ce_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _Z, _P]),
    "alignq_head_ce_bwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "alignq_conv3x3_stem_ws_bytes": (_Z, [_I]),
    "alignq_conv3x3_stem_fwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P]),
    "alignq_conv3x3_stem_bwd_weight": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _Z, _P]),
    "alignq_conv3x3_bwd_weight": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _Z, _P]),
    "alignq_lmmd_ws_bytes": (_Z, [_I]),
    "alignq_lmmd_fwd": (_I, [_P, _I, _I, _P, _F, _I, _F, _P, _P, _P, _Z, _P]),
    "alignq_lmmd_bwd": (_I, [_P, _I, _I, _P, _P, _P, _P, _P]),
    "alignq_sgd_step": (_I, [_P, _P, _P, _I, _L, _F, _F, _I, _F, _P]),
}

_lib = None
_lock = threading.Lock()


class AlignQError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library once; raise loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise AlignQError(
                    f"{LIB_PATH} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(or `make -C alignq_b200/csrc`). alignq_b200 has no CPU or eager fallback.")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
            if lib.alignq_abi_version() != ABI_VERSION:
                raise AlignQError(f"ABI mismatch: library {lib.alignq_abi_version()} vs binding {ABI_VERSION}")
            _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise AlignQError(f"{what} failed ({rc}): {load().alignq_error_string(rc).decode()}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def dev_f32(t: torch.Tensor, what: str) -> torch.Tensor:
    """Validate a kernel operand: CUDA, fp32, contiguous.  No silent host path."""
    if not t.is_cuda:
        raise AlignQError(f"{what}: expected a CUDA tensor, got device {t.device} "
                          "(alignq_b200 runs on sm_100a only; there is no CPU fallback)")
    if t.dtype != torch.float32:
        raise AlignQError(f"{what}: expected float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def dev_f32_dense(t: torch.Tensor, what: str) -> torch.Tensor:
    """Like dev_f32, but keeps any dense layout whose outermost stride is the batch (NCHW-contiguous or
    channels_last): the element-wise kernels are layout-agnostic and the Gram is invariant under a
    permutation of the per-sample features, so no NHWC<->NCHW copy is ever needed."""
    if not t.is_cuda:
        raise AlignQError(f"{what}: expected a CUDA tensor, got device {t.device} "
                          "(alignq_b200 runs on sm_100a only; there is no CPU fallback)")
    if t.dtype != torch.float32:
        raise AlignQError(f"{what}: expected float32, got {t.dtype}")
    if t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last)):
        return t
    return t.contiguous()


def is_dense(t: torch.Tensor) -> bool:
    return t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last))


def phys(t: torch.Tensor) -> torch.Tensor:
    """1-D view of a dense tensor in PHYSICAL memory order (no copy)."""
    return torch.as_strided(t, (t.numel(),), (1,), t.storage_offset())


def like_layout(g: torch.Tensor, ref: torch.Tensor, what: str) -> torch.Tensor:
    """Bring an upstream gradient to the memory layout of the saved activation."""
    if not g.is_cuda or g.dtype != torch.float32:
        raise AlignQError(f"{what}: expected a float32 CUDA tensor")
    if g.stride() == ref.stride():
        return g
    return torch.empty_like(ref).copy_(g)


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def plan_chunks(sizes):
    """Host-side chunk tables for a list of tensor sizes (alignq_wq_plan)."""
    lib = load()
    nseg = len(sizes)
    seg_off = (C.c_int64 * (nseg + 1))()
    acc = 0
    for i, n in enumerate(sizes):
        seg_off[i] = acc
        acc += int(n)
    seg_off[nseg] = acc
    nchunks = lib.alignq_wq_plan(seg_off, nseg, None, None)
    if nchunks < 0:
        check(int(nchunks), "alignq_wq_plan")
    chunk_seg = (C.c_int32 * max(int(nchunks), 1))()
    seg_chunk0 = (C.c_int32 * (nseg + 1))()
    lib.alignq_wq_plan(seg_off, nseg, chunk_seg, seg_chunk0)
    return list(seg_off), list(chunk_seg)[: int(nchunks)], list(seg_chunk0), int(nchunks)
