"""ncu driver stub: the conv kernel VARIANTS the ResNet-20 step launches -- forward with the BN statistics epilogue and data
gradient with the bn-act reduce epilogue at C = 16, fp16 single-term weight gradient at C = 16 / 32, data gradient at C = 32."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignq_b200 import _lib as L
lib = L.load()
cl = lambda t: t.contiguous(memory_format=torch.channels_last)
st = torch.cuda.current_stream().cuda_stream
for rep in range(2):
    for C, H in ((16, 32), (32, 16)):
        N = 128
        x, gy, res = (cl(torch.randn(N, C, H, H, device="cuda")) for _ in range(3))
        y = torch.empty_like(x); gx = torch.empty_like(x)
        w = cl(torch.randn(C, C, 3, 3, device="cuda") * 0.1); gw = torch.empty_like(w)
        ws = torch.empty(int(lib.alignq_conv3x3_ws_bytes(C)), dtype=torch.uint8, device="cuda")
        mean, invstd, g, b, rm, rv, gg, gb = (torch.ones(C, device="cuda") for _ in range(8))
        bws = torch.zeros(int(lib.alignq_bn_act_ws_doubles(C)), dtype=torch.float64, device="cuda")
        cnt = torch.zeros(2, dtype=torch.int32, device="cuda")
        if C == 16:
            L.check(lib.alignq_conv3x3_fwd_bnstats(x.data_ptr(), w.data_ptr(), y.data_ptr(), N, H, H, C, 0, rm.data_ptr(), rv.data_ptr(), 0.1, 1e-5,
                                                   mean.data_ptr(), invstd.data_ptr(), bws.data_ptr(), cnt.data_ptr(), 0, st), "fwd_bnstats")
            L.check(lib.alignq_conv3x3_bwd_data_bnreduce(gy.data_ptr(), w.data_ptr(), gx.data_ptr(), N, H, H, C, 0, y.data_ptr(), x.data_ptr(),
                                                         res.data_ptr(), mean.data_ptr(), invstd.data_ptr(), g.data_ptr(), b.data_ptr(), 2.0, 1,
                                                         gg.data_ptr(), gb.data_ptr(), bws.data_ptr(), cnt.data_ptr(), st), "bwd_data_bnreduce")
        else:
            L.check(lib.alignq_conv3x3_bwd_data(gy.data_ptr(), w.data_ptr(), gx.data_ptr(), N, H, H, C, 0, st), "bwd_data")
        L.check(lib.alignq_conv3x3_bwd_weight(x.data_ptr(), gy.data_ptr(), gw.data_ptr(), N, H, H, C, 0, 0, ws.data_ptr(), ws.numel(), st), "bwd_weight")
torch.cuda.synchronize()
print("ok")
