"""tcgen05 ADMM backward: GPU-side timing vs the FFMA backward (graph replay)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignq_b200 import _lib as L
lib = L.load(); dev = "cuda"
def graph_time(fn, reps=5, iters=5):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (iters * reps) * 1e-3
torch.manual_seed(0)
out = {}
for B, Fd in [(128, 16384), (128, 8192), (128, 4096), (28, 802816), (128, 1 << 20)]:
    x = torch.randn(B, Fd, device=dev); gy = torch.randn(B, Fd, device=dev); gx = torch.empty_like(x)
    dL = torch.randn(B, B, device=dev) * 1e-4; gl = torch.ones(1, device=dev)
    ws = torch.empty(int(lib.alignq_gram_ws_bytes(B, Fd)), dtype=torch.uint8, device=dev)
    ref = None
    for mode, mid in (("fp32", 0), ("tf32x3", 1), ("bf16", 2)):
        call = lambda: L.check(lib.alignq_act_admm_bwd(x.data_ptr(), gy.data_ptr(), dL.data_ptr(), gl.data_ptr(), B, Fd, 8, 2.0, 0.0,
                                                       gx.data_ptr(), ws.data_ptr(), ws.numel(), mid, L.stream_ptr()), "bwd")
        if mode == "fp32" and Fd > 900000: continue
        call(); torch.cuda.synchronize()
        if mode == "fp32": ref = gx.clone()
        err = float((gx - ref).norm() / ref.norm()) if ref is not None else None
        t = graph_time(call)
        out[f"bwd_{mode}_B{B}_F{Fd}"] = {"us": t * 1e6, "gbs_at_12B": 12 * B * Fd / t / 1e9, "relnorm_vs_fp32": err}
        print(f"bwd_{mode}_B{B}_F{Fd}", out[f"bwd_{mode}_B{B}_F{Fd}"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/tcb_probe.json", "w"), indent=1)
