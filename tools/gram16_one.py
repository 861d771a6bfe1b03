import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alignq_b200 import _lib as L
lib = L.load()
B, F = 256, 1 << 20
x = torch.randn(B, F, device="cuda").to(torch.bfloat16)
G = torch.empty(B, B, device="cuda")
ws = torch.empty(int(lib.alignq_gram_bf16_ws_bytes(B)), dtype=torch.uint8, device="cuda")
for _ in range(4):
    L.check(lib.alignq_gram_bf16(x.data_ptr(), B, F, 1, G.data_ptr(), ws.data_ptr(), ws.numel(), L.stream_ptr()), "g")
torch.cuda.synchronize(); print("ok")
