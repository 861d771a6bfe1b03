"""Times the fused act-quant + Gram/ADMM layer (forward incl. loss and dL/dD; backward) on the layer shapes of
BASELINE configs 2 and 5, per gram mode (CUDA-graph replay of the C-ABI call).  Writes gpurun_out/r02_gram_layer_bench.json."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignq_b200 import _lib as L
from tools.tc_probe_util import graph_time
lib = L.load()
dev = "cuda"
out = {}
for B, Fd, eps in [(128, 16384, 0.0), (128, 8192, 0.0), (128, 4096, 0.0), (128, 262144, 0.0), (28, 802816, 1e-5), (28, 100352, 1e-5)]:
    x = torch.randn(B, Fd, device=dev); gy = torch.randn_like(x)
    y, gx = torch.empty_like(x), torch.empty_like(x)
    D = torch.empty(B, B, device=dev); dL = torch.empty(B, B, device=dev); loss = torch.empty((), device=dev)
    gl = torch.ones((), device=dev); Z = torch.rand(B, B, device=dev); U = torch.rand(B, B, device=dev)
    ws = torch.empty(int(lib.alignq_gram_ws_bytes(B, Fd)), dtype=torch.uint8, device=dev)
    for mode, mid in (("tf32x3", 1), ("bf16", 2)):
        fwd = lambda: L.check(lib.alignq_act_admm_fwd(x.data_ptr(), B, Fd, 8, 2.0, eps, Z.data_ptr(), U.data_ptr(), B, 0.2, 0.3,
                              y.data_ptr(), D.data_ptr(), loss.data_ptr(), dL.data_ptr(), ws.data_ptr(), ws.numel(), mid, L.stream_ptr()), "fwd")
        bwd = lambda: L.check(lib.alignq_act_admm_bwd(x.data_ptr(), gy.data_ptr(), dL.data_ptr(), gl.data_ptr(), B, Fd, 8, 2.0, eps,
                              gx.data_ptr(), ws.data_ptr(), ws.numel(), mid, L.stream_ptr()), "bwd")
        tf, tb = graph_time(fwd, reps=5, iters=5), graph_time(bwd, reps=5, iters=5)
        n = x.numel()
        out[f"B{B}_F{Fd}_{mode}"] = {"fwd_us": tf * 1e6, "bwd_us": tb * 1e6, "fwdbwd_gbs_at_20B": 20 * n / (tf + tb) / 1e9}
        print(f"B{B} F{Fd} {mode}: fwd {tf*1e6:.1f} us  bwd {tb*1e6:.1f} us  ({20*n/(tf+tb)/1e9:.0f} GB/s at 20 B/elem)", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r02_gram_layer_bench.json", "w"), indent=1)
