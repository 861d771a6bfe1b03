"""Per-exchange latency of the fused bn-act kernels' global-batch statistics exchange (torchrun, >= 2 GPUs):
forward + backward of bn_act on a tiny and on a ResNet-20-sized tensor, replayed as a CUDA graph of 20 layers,
with local statistics, the NCCL all-reduce between the launches, and the in-kernel NVLink peer exchange."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alignq_b200 as aq  # noqa: E402
from alignq_b200.model.fused import bn_act  # noqa: E402
from alignq_b200.utils import dp_gram  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
dp_gram.configure()
out = {"world": world}
NL = 20
for shape in ((8, 16, 4, 4), (128, 16, 32, 32), (128, 64, 8, 8)):
    for impl in (False, "nccl", "peer"):
        aq.set_args(variant="A", act_range=2, abitW=8, fuse_bn_act=True, method="none", sync_bn=impl)
        torch.manual_seed(0)
        bns = [torch.nn.BatchNorm2d(shape[1]).to(dev).train() for _ in range(NL)]
        q = aq.activation_quantize_fn(8, "second")
        x0 = torch.randn(shape, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)

        def it():
            h = x0
            for bn in bns:
                h = bn_act(bn, q, h, True)
            h.sum().backward()
            x0.grad = None
            for bn in bns:
                bn.weight.grad = bn.bias.grad = None
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                it()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            it()
        for _ in range(5):
            g.replay()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 50 * 1e3
        out[f"{shape} {impl}"] = {"us_per_iteration_of_20_layers_fwd_bwd": us, "us_per_layer_fwd_bwd": us / NL}
        del g
if rank == 0:
    print(json.dumps(out, indent=1))
    json.dump(out, open("gpurun_out/r02_peer_latency_probe.json", "w"), indent=1)
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
