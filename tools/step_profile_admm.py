"""Kernel-time breakdown of one ResNet-56 + ADMM (config 2) QAT step, eager.  Dev tool."""
import os, sys, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alignq_b200 as aq
from alignq_b200.model.resnet import resnet56_quant
from alignq_b200.utils.train import QATStep
from torch.profiler import profile, ProfilerActivity
mode = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
dev = "cuda"
torch.backends.cudnn.benchmark = True
aq.set_args(variant="B", bitW=8, abitW=8, act_range=2, train_batch_size=128, gram_mode=mode, method="ours")
torch.manual_seed(0)
model = resnet56_quant(8, 8, "second").to(dev).train()
step = QATStep(model, channels_last=True)
x = torch.randn(128, 3, 32, 32, device=dev).contiguous(memory_format=torch.channels_last)
t = torch.randint(0, 10, (128,), device=dev)
for _ in range(4): step.step(x, t)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2): step.step(x, t)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if getattr(e.device_type, "name", "") == "CUDA":
        agg[e.name[:84]][0] += 1; agg[e.name[:84]][1] += e.device_time
tot = sum(v[1] for v in agg.values())
print(f"mode={mode} total kernel time per step: {tot/2:.1f} us over {sum(v[0] for v in agg.values())//2} kernels")
for k, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"{d/2:9.1f} us {100*d/tot:5.1f}% x{c//2:4d}  {k}")
