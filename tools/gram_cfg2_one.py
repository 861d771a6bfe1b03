"""ncu driver stub: the fused act-quant + Gram/ADMM layer at config 2's largest shape (B = 128, F = 16384), tf32x3 mode:
forward (gram_tc_kernel with the in-cluster split-K reduction + gram_finish_tc_kernel) and backward."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignq_b200 import _lib as L
lib = L.load()
dev = "cuda"
B, Fd = 128, 16384
xs = [torch.randn(B, Fd, device=dev) for _ in range(4)]
gy = torch.randn(B, Fd, device=dev)
y = torch.empty_like(xs[0]); gx = torch.empty_like(xs[0])
D = torch.empty(B, B, device=dev); dL = torch.empty(B, B, device=dev); loss = torch.empty((), device=dev); gl = torch.ones(1, device=dev)
Z = torch.rand(B, B, device=dev); U = torch.rand(B, B, device=dev)
ws = torch.empty(int(lib.alignq_gram_ws_bytes(B, Fd)), dtype=torch.uint8, device=dev)
for i in range(4):
    x = xs[i]
    L.check(lib.alignq_act_admm_fwd(x.data_ptr(), B, Fd, 8, 2.0, 0.0, Z.data_ptr(), U.data_ptr(), B, 0.2, 0.3, y.data_ptr(), D.data_ptr(),
                                    loss.data_ptr(), dL.data_ptr(), ws.data_ptr(), ws.numel(), 1, L.stream_ptr()), "fused")
    L.check(lib.alignq_act_admm_bwd(x.data_ptr(), gy.data_ptr(), dL.data_ptr(), gl.data_ptr(), B, Fd, 8, 2.0, 0.0, gx.data_ptr(),
                                    ws.data_ptr(), ws.numel(), 1, L.stream_ptr()), "bwd")
torch.cuda.synchronize()
print("ok", float(loss))
