"""tcgen05 Gram bring-up: correctness vs fp64 and GPU-side timing (CUDA-graph replay removes the
Python/ctypes launch overhead) of the three numerics modes."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alignq_b200 as aq
from alignq_b200 import _lib as L
from oracle import alignq_oracle as O
dev = "cuda"
out = {}
lib = L.load()

def graph_time(fn, reps=10, iters=10):
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
MODES = {"fp32": 0, "tf32x3": 1, "bf16": 2}
for B, F in [(128, 4096), (128, 16384), (28, 100352), (128, 262144), (128, 1 << 21)]:
    x = torch.randn(B, F, device=dev)
    ref = O.corr(x[:, : min(F, 262144)].double(), x[:, : min(F, 262144)].double(), 0.0) if F <= 262144 else None
    G = torch.empty(B, B, device=dev)
    ws = torch.empty(int(lib.alignq_gram_ws_bytes(B, F)), dtype=torch.uint8, device=dev)
    for mode, mid in MODES.items():
        if mode == "fp32" and F > 262144: continue
        call = lambda: L.check(lib.alignq_corr_fwd(x.data_ptr(), x.data_ptr(), B, F, 0.0, G.data_ptr(), ws.data_ptr(), ws.numel(), mid, L.stream_ptr()), "corr")
        call(); torch.cuda.synchronize()
        err = float((G.double() - ref).abs().max() / ref.abs().max()) if ref is not None else None
        t = graph_time(call)
        out[f"corr_{mode}_B{B}_F{F}"] = {"err_over_maxG": err, "us": t * 1e6, "tflops": 2 * B * B * F / t / 1e12, "gbs": 4 * B * F / t / 1e9}
        print(f"corr_{mode}_B{B}_F{F}", out[f"corr_{mode}_B{B}_F{F}"], flush=True)
for B, shape in [(128, (16, 32, 32)), (128, (32, 16, 16)), (128, (64, 8, 8)), (28, (256, 56, 56)), (128, (256, 64, 64))]:
    x = torch.randn(B, *shape, device=dev)
    n = x.numel(); Fd = n // B
    y = torch.empty_like(x); D = torch.empty(B, B, device=dev); dL = torch.empty(B, B, device=dev); loss = torch.empty((), device=dev)
    Z = torch.rand(B, B, device=dev); U = torch.rand(B, B, device=dev)
    ws = torch.empty(int(lib.alignq_gram_ws_bytes(B, Fd)), dtype=torch.uint8, device=dev)
    for mode, mid in MODES.items():
        call = lambda: L.check(lib.alignq_act_admm_fwd(x.data_ptr(), B, Fd, 8, 2.0, 0.0, Z.data_ptr(), U.data_ptr(), B, 0.2, 0.3,
                               y.data_ptr(), D.data_ptr(), loss.data_ptr(), dL.data_ptr(), ws.data_ptr(), ws.numel(), mid, L.stream_ptr()), "fused")
        t = graph_time(call, reps=5, iters=5)
        out[f"fused_fwd_{mode}_B{B}_F{Fd}"] = {"us": t * 1e6, "gbs_at_8B": 8 * n / t / 1e9, "tflops": 4 * B * B * Fd / t / 1e12}
        print(f"fused_fwd_{mode}_B{B}_F{Fd}", out[f"fused_fwd_{mode}_B{B}_F{Fd}"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/tc_probe.json", "w"), indent=1)
