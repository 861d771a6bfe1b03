"""Dev aid: where does the time of one conv3x3 forward / data-gradient launch go?  Loads tools/libalignq_conv_trace.so
(csrc/conv_tc.cu + abi.cu built with -DALIGNQ_CONV_TRACE: thread 0 of every CTA stamps %globaltimer at the phase
boundaries) and prints, per phase, the mean / max over the CTAs and the spread of the CTAs' start times.

  make -C alignq_b200/csrc trace        # -> tools/libalignq_conv_trace.so (git-ignored; travels to the GPU box)
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
lib = C.CDLL(os.path.join(HERE, "libalignq_conv_trace.so"))
P, I = C.c_void_p, C.c_int
for name in ("alignq_conv3x3_fwd", "alignq_conv3x3_bwd_data"):
    getattr(lib, name).argtypes = [P, P, P, I, I, I, I, I, P]
    getattr(lib, name).restype = I
lib.alignq_conv_trace_read.argtypes = [P, C.c_size_t]
lib.alignq_conv3x3_bwd_weight.argtypes = [P, P, P, I, I, I, I, I, I, P, C.c_size_t, P]
lib.alignq_conv3x3_bwd_weight.restype = I
lib.alignq_conv3x3_ws_bytes.argtypes = [I]
lib.alignq_conv3x3_ws_bytes.restype = C.c_size_t
SLOTS, CTAS = 16, 1024
NAMES = {0: "entry", 1: "prologue done (weights / gy scale)", 2: "tile0 deposited", 3: "tile0 MMAs issued", 4: "tile0 MMAs done",
         5: "tile0 epilogue done", 6: "tile1 deposited", 7: "tile1 MMAs issued", 8: "tile1 MMAs done", 9: "tile1 epilogue done",
         13: "loop done", 14: "exit"}


def run(fn, shape, flush):
    N, Cc, H, W = shape
    x = torch.randn(shape, device="cuda").contiguous(memory_format=torch.channels_last)
    w = torch.randn(Cc, Cc, 3, 3, device="cuda").contiguous(memory_format=torch.channels_last)
    y = torch.empty_like(x)
    junk = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    ws = torch.empty(int(lib.alignq_conv3x3_ws_bytes(Cc)), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        if flush:
            junk.zero_()
        torch.cuda.synchronize()
        assert lib.alignq_conv_trace_reset() == 0
        if fn == "alignq_conv3x3_bwd_weight":       # x, gy -> gw (w is the output here)
            rc = lib.alignq_conv3x3_bwd_weight(x.data_ptr(), y.normal_().data_ptr(), w.data_ptr(), N, H, W, Cc, 0, 0, ws.data_ptr(), ws.numel(), s)
        else:
            rc = getattr(lib, fn)(x.data_ptr(), w.data_ptr(), y.data_ptr(), N, H, W, Cc, 0, s)
        assert rc == 0, rc
    torch.cuda.synchronize()
    buf = np.zeros(CTAS * SLOTS, dtype=np.uint64)
    assert lib.alignq_conv_trace_read(buf.ctypes.data, buf.nbytes) == 0
    t = buf.reshape(CTAS, SLOTS).astype(np.int64)
    live = t[:, 0] > 0
    t = t[live]
    t0 = t[:, 0].min()
    print(f"== {fn} {shape} flush={flush}: {live.sum()} CTAs on {len(set(t[:, 15]))} SMs; kernel span "
          f"{(t[:, 14].max() - t0) / 1e3:.2f} us; CTA start spread {(t[:, 0].max() - t0) / 1e3:.2f} us")
    prev = 0
    for k in sorted(NAMES):
        ok = t[:, k] > 0
        if not ok.any() or k == 0:
            continue
        d = (t[ok, k] - t[ok, prev]) / 1e3
        a = (t[ok, k] - t0) / 1e3
        print(f"  {NAMES[k]:28s} +{d.mean():6.2f} us (max {d.max():6.2f})   at {a.mean():6.2f} us (max {a.max():6.2f})   [{ok.sum()} CTAs]")
        prev = k


if __name__ == "__main__":
    shapes = [(128, 16, 32, 32), (128, 32, 16, 16), (128, 64, 8, 8)]
    for shape in shapes:
        for fn in (sys.argv[1:] or ["alignq_conv3x3_fwd", "alignq_conv3x3_bwd_data", "alignq_conv3x3_bwd_weight"]):
            run(fn, shape, True)
    run("alignq_conv3x3_fwd", shapes[0], False)
