"""Small-batch (B <= 32) tcgen05 Gram bring-up: correctness vs fp64 + GPU-side timing (CUDA-graph replay)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alignq_b200 as aq
from alignq_b200 import _lib as L
from oracle import alignq_oracle as O
from tools.tc_probe_util import graph_time
dev = "cuda"
out = {}
lib = L.load()
torch.manual_seed(0)
MODES = {"fp32": 0, "tf32x3": 1, "bf16": 2}
for B, F in [(28, 3000), (28, 100352), (32, 65536), (8, 4096), (28, 802816)]:
    x = torch.randn(B, F, device=dev) * 1.3 + 0.2
    ref = O.corr(x.double(), x.double(), 1e-5)
    G = torch.empty(B, B, device=dev)
    ws = torch.empty(int(lib.alignq_gram_ws_bytes(B, F)), dtype=torch.uint8, device=dev)
    for mode, mid in MODES.items():
        call = lambda: L.check(lib.alignq_corr_fwd(x.data_ptr(), x.data_ptr(), B, F, 1e-5, G.data_ptr(), ws.data_ptr(), ws.numel(), mid, L.stream_ptr()), "corr")
        call(); torch.cuda.synchronize()
        err = float((G.double() - ref).abs().max() / ref.abs().max())
        t = graph_time(call)
        out[f"corr_{mode}_B{B}_F{F}"] = {"err_over_maxG": err, "us": t * 1e6, "gbs": 4 * B * F / t / 1e9}
        print(f"corr_{mode}_B{B}_F{F}", out[f"corr_{mode}_B{B}_F{F}"], flush=True)
for B, shape in [(28, (256, 56, 56)), (28, (512, 28, 28)), (28, (1024, 14, 14)), (28, (2048, 7, 7)), (28, (64, 112, 112))]:
    x = torch.randn(B, *shape, device=dev)
    n = x.numel(); Fd = n // B
    y = torch.empty_like(x); D = torch.empty(B, B, device=dev); dL = torch.empty(B, B, device=dev); loss = torch.empty((), device=dev)
    gy = torch.randn_like(x); gx = torch.empty_like(x); gl = torch.ones((), device=dev)
    Z = torch.rand(B, B, device=dev); U = torch.rand(B, B, device=dev)
    ws = torch.empty(int(lib.alignq_gram_ws_bytes(B, Fd)), dtype=torch.uint8, device=dev)
    res = {}
    for mode, mid in MODES.items():
        call = lambda: L.check(lib.alignq_act_admm_fwd(x.data_ptr(), B, Fd, 8, 2.0, 1e-5, Z.data_ptr(), U.data_ptr(), B, 0.2, 0.3,
                               y.data_ptr(), D.data_ptr(), loss.data_ptr(), dL.data_ptr(), ws.data_ptr(), ws.numel(), mid, L.stream_ptr()), "fused")
        t = graph_time(call, reps=5, iters=5)
        callb = lambda: L.check(lib.alignq_act_admm_bwd(x.data_ptr(), gy.data_ptr(), dL.data_ptr(), gl.data_ptr(), B, Fd, 8, 2.0, 1e-5,
                                gx.data_ptr(), ws.data_ptr(), ws.numel(), mid, L.stream_ptr()), "bwd")
        tb = graph_time(callb, reps=5, iters=5)
        res[mode] = (y.clone(), D.clone(), gx.clone())
        k = f"fused_{mode}_B{B}_F{Fd}"
        out[k] = {"fwd_us": t * 1e6, "fwd_gbs_at_8B": 8 * n / t / 1e9, "bwd_us": tb * 1e6, "bwd_gbs_at_12B": 12 * n / tb / 1e9}
        if mode != "fp32":
            out[k]["y_equal"] = bool(torch.equal(res[mode][0], res["fp32"][0]))
            out[k]["dD_abs"] = float((res[mode][1] - res["fp32"][1]).abs().max())
            out[k]["gx_relnorm"] = float((res[mode][2] - res["fp32"][2]).norm() / res["fp32"][2].norm())
        print(k, out[k], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/tcs_probe.json", "w"), indent=1)
