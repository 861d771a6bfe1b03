"""Kernel-time breakdown of one QAT step with torch.profiler (CUPTI), eager (no graph).  Dev tool."""
import os, sys, json, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alignq_b200 as aq
from alignq_b200.model.resnet import resnet20_quant
from alignq_b200.utils.train import QATStep
from torch.profiler import profile, ProfilerActivity
fuse = "--no-fuse" not in sys.argv
dev = "cuda"
torch.backends.cudnn.benchmark = True
aq.set_args(variant="A", bitW=8, abitW=8, act_range=2, train_batch_size=128, fuse_bn_act=fuse)
torch.manual_seed(0)
model = resnet20_quant(8, 8, "second").to(dev).train()
step = QATStep(model, channels_last=True)
x = torch.randn(128, 3, 32, 32, device=dev).contiguous(memory_format=torch.channels_last)
t = torch.randint(0, 10, (128,), device=dev)
for _ in range(5): step.step(x, t)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step.step(x, t)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type.name == "CUDA" if hasattr(e.device_type, "name") else False:
        agg[e.name[:80]][0] += 1; agg[e.name[:80]][1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in agg.values())
print(f"fuse={fuse} total kernel time per step: {tot/3:.1f} us over {sum(v[0] for v in agg.values())//3} kernels")
for k, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{d/3:9.1f} us {100*d/tot:5.1f}% x{c//3:4d}  {k}")
