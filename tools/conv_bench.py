"""Times the tcgen05 3x3 convolution kernels against cuDNN (TF32 on, channels_last, cudnn.benchmark) on the layer shapes
of resnet20_quant at batch 128: forward, data gradient, weight gradient; CUDA events over graph-replayed batches of
launches on rotating buffers (so the 126 MB L2 does not hold one layer's tensors across iterations)."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignq_b200 import _lib as L  # noqa: E402

lib = L.load()
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = True
dev = "cuda"
cl = lambda t: t.contiguous(memory_format=torch.channels_last)
out = {}
NB = 12                                                       # rotating buffer sets: 12 x (x, y, gy, gx) of 8 MB = 400 MB


def timeit(fn, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(NB):
            fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(NB):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * NB) * 1e3


for (N, H, W, C) in [(128, 32, 32, 16), (128, 16, 16, 32), (128, 8, 8, 64)]:
    xs = [cl(torch.randn(N, C, H, W, device=dev)) for _ in range(NB)]
    gys = [cl(torch.randn(N, C, H, W, device=dev)) for _ in range(NB)]
    ys = [torch.empty_like(x) for x in xs]
    w = cl(torch.randn(C, C, 3, 3, device=dev) * 0.1)
    gw = torch.empty_like(w)
    ws = torch.empty(int(lib.alignq_conv3x3_ws_bytes(C)), dtype=torch.uint8, device=dev)
    st = lambda: torch.cuda.current_stream().cuda_stream
    row = {}
    for mode, name in ((0, "tf32"), (1, "tf32x3")):
        if C == 64 and mode == 1:
            continue
        row[f"own_{name}_fwd_us"] = timeit(lambda i: L.check(lib.alignq_conv3x3_fwd(xs[i].data_ptr(), w.data_ptr(), ys[i].data_ptr(), N, H, W, C, mode, st()), "f"))
        row[f"own_{name}_dgrad_us"] = timeit(lambda i: L.check(lib.alignq_conv3x3_bwd_data(gys[i].data_ptr(), w.data_ptr(), ys[i].data_ptr(), N, H, W, C, mode, st()), "d"))
        row[f"own_{name}_wgrad_us"] = timeit(lambda i: L.check(lib.alignq_conv3x3_bwd_weight(xs[i].data_ptr(), gys[i].data_ptr(), gw.data_ptr(), N, H, W, C, mode, 0, ws.data_ptr(), ws.numel(), st()), "w"))
    row["cudnn_tf32_fwd_us"] = timeit(lambda i: F.conv2d(xs[i], w, None, 1, 1))
    row["cudnn_tf32_dgrad_us"] = timeit(lambda i: torch.ops.aten.convolution_backward(gys[i], xs[i], w, None, (1, 1), (1, 1), (1, 1), False, (0, 0), 1, (True, False, False)))
    row["cudnn_tf32_wgrad_us"] = timeit(lambda i: torch.ops.aten.convolution_backward(gys[i], xs[i], w, None, (1, 1), (1, 1), (1, 1), False, (0, 0), 1, (False, True, False)))
    row["bytes_in_out_MB"] = 2 * N * H * W * C * 4 / 1e6
    out[f"N{N}_{H}x{W}_C{C}"] = row
    print(f"N{N} {H}x{W} C{C}", json.dumps({k: round(v, 2) for k, v in row.items()}), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r02_conv_bench.json", "w"), indent=1)
