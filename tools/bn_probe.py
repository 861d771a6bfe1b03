"""Fused BN -> act-quant -> ReLU kernels on the BN shapes of MobileNet-v2 (B=256) and DenseNet-40 (B=128):
GPU time per launch pair (CUDA-graph replay) and GB/s at the algorithmic 12 B/elem forward, 28 B/elem backward."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignq_b200 import _lib as L
from tools.tc_probe_util import graph_time
lib = L.load()
dev = "cuda"
SHAPES = [(256, 16, 32, 32), (256, 24, 32, 32), (256, 32, 32, 32), (256, 96, 32, 32), (256, 144, 32, 32), (256, 144, 16, 16),
          (256, 192, 16, 16), (256, 32, 16, 16), (256, 384, 8, 8), (256, 576, 8, 8), (256, 64, 8, 8), (256, 960, 4, 4),
          (256, 160, 4, 4), (128, 24, 32, 32), (128, 84, 32, 32), (128, 156, 32, 32), (128, 168, 16, 16), (128, 300, 16, 16),
          (128, 312, 8, 8), (128, 456, 8, 8), (128, 16, 32, 32), (128, 64, 8, 8)]
out = {}
tot_f = tot_b = 0.0
for (B, C, H, W) in SHAPES:
    rows = B * H * W
    x = torch.randn(rows, C, device=dev); gy = torch.randn(rows, C, device=dev)
    y = torch.empty_like(x); gx = torch.empty_like(x)
    g = torch.rand(C, device=dev) + 0.5; b = torch.randn(C, device=dev) * 0.1
    rm = torch.zeros(C, device=dev); rv = torch.ones(C, device=dev)
    mean = torch.empty(C, device=dev); invstd = torch.empty(C, device=dev)
    gg = torch.empty(C, device=dev); gb = torch.empty(C, device=dev)
    ws = torch.zeros(int(lib.alignq_bn_act_ws_doubles(C)), dtype=torch.float64, device=dev)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    fwd = lambda: L.check(lib.alignq_bn_act_fwd(x.data_ptr(), rows, C, g.data_ptr(), b.data_ptr(), rm.data_ptr(), rv.data_ptr(), 0.1, 1e-5, 1,
                          8, 2.0, 0, 1, 0, y.data_ptr(), mean.data_ptr(), invstd.data_ptr(), ws.data_ptr(), counter.data_ptr(), 0, L.stream_ptr()), "fwd")
    bwd = lambda: L.check(lib.alignq_bn_act_bwd(x.data_ptr(), y.data_ptr(), gy.data_ptr(), rows, C, g.data_ptr(), b.data_ptr(), mean.data_ptr(),
                          invstd.data_ptr(), 1, 8, 2.0, 0, 1, gx.data_ptr(), 0, gg.data_ptr(), gb.data_ptr(), ws.data_ptr(), counter.data_ptr(),
                          L.stream_ptr()), "bwd")
    tf = graph_time(fwd, reps=5, iters=5); tb = graph_time(bwd, reps=5, iters=5)
    n = rows * C
    rec = {"MB": 4 * n / 1e6, "fwd_us": tf * 1e6, "bwd_us": tb * 1e6, "fwd_gbs_at_12B": 12 * n / tf / 1e9, "bwd_gbs_at_28B": 28 * n / tb / 1e9}
    out[f"{B}x{C}x{H}x{W}"] = rec
    print(f"{B}x{C}x{H}x{W}", {k: round(v, 1) for k, v in rec.items()}, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bn_probe.json", "w"), indent=1)
