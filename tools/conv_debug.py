import sys, os, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignq_b200 import _lib as L
lib = L.load()
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(0)
N, H, W, C = 2, 8, 8, 16
cl = lambda t: t.contiguous(memory_format=torch.channels_last)
x = cl(torch.randn(N, C, H, W, device="cuda")); gy = cl(torch.randn(N, C, H, W, device="cuda"))
w = cl(torch.randn(C, C, 3, 3, device="cuda"))
ref = torch.nn.grad.conv2d_weight(x, w.shape, gy, padding=1)
for mode in (0, 1):
    gw = cl(torch.full((C, C, 3, 3), 7.0, device="cuda"))
    ws = torch.full((int(lib.alignq_conv3x3_ws_bytes(C)) // 4,), 3.0, device="cuda")
    rc = lib.alignq_conv3x3_bwd_weight(x.data_ptr(), gy.data_ptr(), gw.data_ptr(), N, H, W, C, mode, 0, ws.data_ptr(), ws.numel() * 4, L.stream_ptr())
    torch.cuda.synchronize()
    print("mode", mode, "rc", rc, "gw absmax", float(gw.abs().max()), "nonzero", int((gw != 0).sum()), "==7", int((gw == 7).sum()),
          "ws changed", int((ws != 3.0).sum()), "ws absmax(changed)", float(ws[ws != 3.0].abs().max()) if int((ws != 3.0).sum()) else None)
    print(" ref absmax", float(ref.abs().max()), "err", float((gw - ref).abs().max()))
    part = ws[: C * 9 * C].view(C, 9, C)            # [co][tap][ci]
    refp = ref.permute(0, 2, 3, 1).reshape(C, 9, C)
    print(" partial vs ref err", float((part - refp).abs().max()), " part[0,0,:4]", part[0, 0, :4].tolist(), "ref", refp[0, 0, :4].tolist())
    print(" part[0,4,:4]", part[0, 4, :4].tolist(), "ref", refp[0, 4, :4].tolist())
