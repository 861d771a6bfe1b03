"""Multi-GPU parity evidence (NCCL over NVLink; run under torchrun on >= 2 GPUs of one box):

  1. feature-sharded global-batch Gram: every rank holds a batch shard, an all-to-all gives it all rows of its
     feature slice, the tcgen05 corr kernel forms the partial Gram, ONE all-reduce sums the partials
     (utils/sharding.py) -- compared with the single-device corr of the full [B_global, F] matrix;
  2. (sum, sum-of-squares) all-reduce: combine_moments over a sharded tensor vs torch.mean / torch.std;
  3. data-parallel QAT step (batch sharded, SyncBatchNorm, one gradient all-reduce) vs the single-device step on
     the whole batch: first-iteration loss and parameters after the step.

Writes gpurun_out/dp_parity.json on rank 0.  Evidence tool; the CPU/gloo twins of 1 and 2 are in
tests/test_sharding_gloo.py."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alignq_b200 as aq  # noqa: E402
from alignq_b200.model.resnet import resnet20_quant  # noqa: E402
from alignq_b200.utils import sharding as S  # noqa: E402
from alignq_b200.utils.train import QATStep  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
out = {"world": world, "backend": "nccl"}

# ---- 1. feature-sharded Gram ------------------------------------------------------------------------------------
for Bg, F, eps in [(world * 28, 100352, 1e-5), (128, 16384, 0.0)]:
    if Bg % world or Bg > 128:
        continue
    torch.manual_seed(7)                                         # same global matrix on every rank
    xg = torch.randn(Bg, F, device=dev) * 1.3 + 0.2
    mine = S.shard_batch(xg, rank, world).contiguous()          # what this rank would hold in data-parallel training
    bounds = [S.feature_slice(F, r, world) for r in range(world)]
    send = [mine[:, b0:b1].contiguous() for (b0, b1) in bounds]
    f0, f1 = bounds[rank]
    recv = [torch.empty(S.shard_batch(xg, r, world).shape[0], f1 - f0, device=dev) for r in range(world)]
    dist.all_to_all(recv, send)                                   # rows of every rank, my feature slice
    cols = torch.cat(recv, dim=0).contiguous()
    for mode, tol in (("tf32x3", 1e-5), ("fp32", 1e-5)):
        aq.set_args(gram_mode=mode)
        G = aq.corr(cols, cols, eps)
        S.allreduce_gram_partial_(G, f1 - f0, F)
        ref = None
        xs = ((xg.double() - xg.double().mean(0)) / (xg.double().std(0) + eps))
        ref = xs @ xs.t() / F
        err = float((G.double() - ref).abs().max() / ref.abs().max())
        out[f"gram_feature_sharded_{mode}_B{Bg}_F{F}"] = {"err_over_maxG_vs_fp64_single_device": err, "tol": tol, "ok": err <= tol}
        assert err <= tol, (mode, Bg, F, err)

# ---- 2. (sum, sumsq) all-reduce ----------------------------------------------------------------------------------
torch.manual_seed(11)
w = torch.randn(64 * 64 * 9, device=dev) * 0.05 + 0.01
part = S.shard_batch(w, rank, world)
mean, std = S.combine_moments(part.double().sum(), (part.double() ** 2).sum(), torch.tensor(float(part.numel()), device=dev))
e_m = abs(float(mean) - float(w.mean())) / abs(float(w.mean()))
e_s = abs(float(std) - float(w.std())) / float(w.std())
out["combine_moments"] = {"mean_rel_err": e_m, "std_rel_err": e_s, "ok": e_m <= 1e-5 and e_s <= 1e-6}
assert e_m <= 1e-5 and e_s <= 1e-6

# ---- 3. data-parallel step vs single-device step --------------------------------------------------------------------
Bglob = 128
aq.reset_args()
aq.set_args(variant="A", bitW=8, abitW=8, act_range=2, lam=1.0, lam2=4.0, train_batch_size=Bglob, fuse_bn_act=False)
torch.manual_seed(0)
single = resnet20_quant(8, 8, "second").to(dev).train()
torch.manual_seed(0)
shard = torch.nn.SyncBatchNorm.convert_sync_batchnorm(resnet20_quant(8, 8, "second")).to(dev).train()
g = torch.Generator().manual_seed(5)
x = torch.randn(Bglob, 3, 32, 32, generator=g).to(dev)
t = torch.randint(0, 10, (Bglob,), generator=g).to(dev)
s1 = QATStep(single, lr=0.04, momentum=0.9, weight_decay=1e-4)
sN = QATStep(shard, lr=0.04, momentum=0.9, weight_decay=1e-4, world_size=world)
l1 = float(s1.step(x, t))
lN = sN.step(S.shard_batch(x, rank, world).contiguous(), S.shard_batch(t, rank, world).contiguous()).detach().clone()
dist.all_reduce(lN)
lN = float(lN) / world
num = sum(float((p.detach().double() - q.detach().double()).pow(2).sum()) for p, q in zip(shard.parameters(), single.parameters()))
den = sum(float(q.detach().double().pow(2).sum()) for q in single.parameters())
out["dp_step_vs_single_device"] = {"loss_single": l1, "loss_dp_mean": lN, "loss_rel_err": abs(lN - l1) / abs(l1),
                                   "params_after_step_relnorm": (num / den) ** 0.5,
                                   "note": "SyncBatchNorm + one gradient all-reduce; differences come from rounding ties that flip "
                                           "when the BN statistics differ in the last bits (see profiles/r01_chaos_band.json)"}
assert abs(lN - l1) <= 1e-4 * abs(l1)

if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/dp_parity.json", "w"), indent=1)
    print(json.dumps(out, indent=1))
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
