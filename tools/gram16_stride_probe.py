"""Does the bf16 Gram kernel's HBM rate depend on the row stride (power-of-two strides -> DRAM channel camping)?"""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignq_b200 import _lib as L
from tools.tc_probe_util import graph_time
lib = L.load(); dev = "cuda"
out = {}
for B in (256, 128):
    for F in (1 << 20, (1 << 20) + 4096, (1 << 20) + 64 * 37, 1000000, 802816, 1 << 22, (1 << 22) + 64 * 37):
        x = torch.randn(B, F, device=dev).to(torch.bfloat16)
        G = torch.empty(B, B, device=dev)
        ws = torch.empty(int(lib.alignq_gram_bf16_ws_bytes(B)), dtype=torch.uint8, device=dev)
        t = graph_time(lambda: L.check(lib.alignq_gram_bf16(x.data_ptr(), B, F, 1, G.data_ptr(), ws.data_ptr(), ws.numel(), L.stream_ptr()), "g16"), reps=5, iters=5)
        out[f"B{B}_F{F}"] = {"us": t * 1e6, "tflops": 2 * B * B * F / t / 1e12, "hbm_gbs": 2 * B * F / t / 1e9}
        print(f"B{B}_F{F}", out[f"B{B}_F{F}"], flush=True)
        del x
json.dump(out, open("gpurun_out/gram16_stride_probe.json", "w"), indent=1)
