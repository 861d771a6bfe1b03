"""Minimal driver for ncu: fused BN->act-quant->ReLU forward + backward on MobileNet-v2's largest BN shape."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignq_b200 import _lib as L
lib = L.load()
dev = "cuda"
B, C, H, W = (int(a) for a in (sys.argv[1:5] if len(sys.argv) >= 5 else (256, 144, 32, 32)))
rows = B * H * W
x = torch.randn(rows, C, device=dev); gy = torch.randn(rows, C, device=dev)
y = torch.empty_like(x); gx = torch.empty_like(x)
g = torch.rand(C, device=dev) + 0.5; b = torch.randn(C, device=dev) * 0.1
rm = torch.zeros(C, device=dev); rv = torch.ones(C, device=dev)
mean = torch.empty(C, device=dev); invstd = torch.empty(C, device=dev)
gg = torch.empty(C, device=dev); gb = torch.empty(C, device=dev)
ws = torch.zeros(int(lib.alignq_bn_act_ws_doubles(C)), dtype=torch.float64, device=dev)
counter = torch.zeros(2, dtype=torch.int32, device=dev)
for _ in range(2):
    L.check(lib.alignq_bn_act_fwd(x.data_ptr(), rows, C, g.data_ptr(), b.data_ptr(), rm.data_ptr(), rv.data_ptr(), 0.1, 1e-5, 1,
            8, 2.0, 0, 1, 0, y.data_ptr(), mean.data_ptr(), invstd.data_ptr(), ws.data_ptr(), counter.data_ptr(), 0, L.stream_ptr()), "fwd")
    L.check(lib.alignq_bn_act_bwd(x.data_ptr(), y.data_ptr(), gy.data_ptr(), rows, C, g.data_ptr(), b.data_ptr(), mean.data_ptr(),
            invstd.data_ptr(), 1, 8, 2.0, 0, 1, gx.data_ptr(), 0, gg.data_ptr(), gb.data_ptr(), ws.data_ptr(), counter.data_ptr(),
            L.stream_ptr()), "bwd")
torch.cuda.synchronize()
print("ok", float(gx.abs().mean()))
