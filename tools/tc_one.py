"""Minimal driver for ncu: a few launches of the tcgen05 Gram kernels."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignq_b200 import _lib as L
lib = L.load()
dev = "cuda"
B, Fd = 128, 262144
x = torch.randn(B, Fd, device=dev)
y = torch.empty_like(x); D = torch.empty(B, B, device=dev); dL = torch.empty(B, B, device=dev); loss = torch.empty((), device=dev)
G = torch.empty(B, B, device=dev)
Z = torch.rand(B, B, device=dev); U = torch.rand(B, B, device=dev)
ws = torch.empty(int(lib.alignq_gram_ws_bytes(B, Fd)), dtype=torch.uint8, device=dev)
for rep in range(3):
    for mid in (1, 2):
        L.check(lib.alignq_act_admm_fwd(x.data_ptr(), B, Fd, 8, 2.0, 0.0, Z.data_ptr(), U.data_ptr(), B, 0.2, 0.3, y.data_ptr(), D.data_ptr(),
                                        loss.data_ptr(), dL.data_ptr(), ws.data_ptr(), ws.numel(), mid, L.stream_ptr()), "fused")
        L.check(lib.alignq_corr_fwd(x.data_ptr(), x.data_ptr(), B, Fd, 0.0, G.data_ptr(), ws.data_ptr(), ws.numel(), mid, L.stream_ptr()), "corr")
torch.cuda.synchronize()
print("ok", float(loss))
