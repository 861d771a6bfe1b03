"""ncu driver stub: a few launches of the tcgen05 conv kernels on the stage-1 shape of resnet20_quant (B=128, 32x32, C=16)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignq_b200 import _lib as L
lib = L.load()
N, H, W, C = 128, 32, 32, int(os.environ.get("CONV_C", "16"))
if C == 32: H = W = 16
if C == 64: H = W = 8
mode = int(os.environ.get("CONV_MODE", "0"))
cl = lambda t: t.contiguous(memory_format=torch.channels_last)
xs = [cl(torch.randn(N, C, H, W, device="cuda")) for _ in range(4)]
gys = [cl(torch.randn(N, C, H, W, device="cuda")) for _ in range(4)]
y = torch.empty_like(xs[0]); w = cl(torch.randn(C, C, 3, 3, device="cuda") * 0.1); gw = torch.empty_like(w)
ws = torch.empty(int(lib.alignq_conv3x3_ws_bytes(C)), dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for i in range(4):
    L.check(lib.alignq_conv3x3_fwd(xs[i].data_ptr(), w.data_ptr(), y.data_ptr(), N, H, W, C, mode, st), "f")
    L.check(lib.alignq_conv3x3_bwd_data(gys[i].data_ptr(), w.data_ptr(), y.data_ptr(), N, H, W, C, mode, st), "d")
    L.check(lib.alignq_conv3x3_bwd_weight(xs[i].data_ptr(), gys[i].data_ptr(), gw.data_ptr(), N, H, W, C, mode, 0, ws.data_ptr(), ws.numel(), st), "w")
torch.cuda.synchronize()
print("ok")
