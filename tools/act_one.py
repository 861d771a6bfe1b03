"""Minimal driver for ncu --set full: the CDF-quantizer kernel pair on the bench's roofline input."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignq_b200 import _lib as L
lib = L.load()
n = 256 * (1 << 20)
x = torch.randn(n, device="cuda"); gy = torch.randn(n, device="cuda")
y = torch.empty_like(x); gx = torch.empty_like(x)
for _ in range(3):
    L.check(lib.alignq_act_fwd(x.data_ptr(), y.data_ptr(), 0, n, 8, 2.0, 0, 0, L.stream_ptr()), "fwd")
    L.check(lib.alignq_act_bwd(x.data_ptr(), gy.data_ptr(), gx.data_ptr(), n, 8, 2.0, 0, 0, L.stream_ptr()), "bwd")
torch.cuda.synchronize()
print("ok")
