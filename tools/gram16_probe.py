"""bf16 TMA + tcgen05 Gram micro-kernel: correctness vs fp64 and GPU-side timing (graph replay)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignq_b200 import _lib as L
lib = L.load()
dev = "cuda"
out = {}
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
for B, F in [(256, 4096), (256, 65536), (128, 65536), (100, 8192), (256, 1 << 20), (128, 1 << 20), (256, 1 << 22)]:
    x = torch.randn(B, F, device=dev).to(torch.bfloat16)
    G = torch.empty(B, B, device=dev)
    ws = torch.empty(int(lib.alignq_gram_bf16_ws_bytes(B)), dtype=torch.uint8, device=dev)
    call = lambda: L.check(lib.alignq_gram_bf16(x.data_ptr(), B, F, 1, G.data_ptr(), ws.data_ptr(), ws.numel(), L.stream_ptr()), "gram_bf16")
    call(); torch.cuda.synchronize()
    err = None
    if F <= 65536:
        ref = (x.double() @ x.double().t()) / F
        err = float((G.double() - ref).abs().max() / ref.abs().max())
    else:
        xs = x[:, :65536]
        # spot check: G * F restricted is not separable; compare against torch bf16 matmul with fp32 accumulate
        ref = (x.float() @ x.float().t()) / F
        err = float((G - ref).abs().max() / ref.abs().max())
    t = graph_time(call)
    tm = graph_time(lambda: torch.matmul(x, x.t()))
    out[f"gram_bf16_B{B}_F{F}"] = {"err_over_maxG": err, "us": t * 1e6, "tflops": 2 * B * B * F / t / 1e12,
                                   "hbm_gbs": 2 * B * F / t / 1e9, "torch_matmul_bf16_us": tm * 1e6,
                                   "torch_matmul_tflops": 2 * B * B * F / tm / 1e12}
    print(f"gram_bf16_B{B}_F{F}", out[f"gram_bf16_B{B}_F{F}"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gram16_probe.json", "w"), indent=1)
