"""Multi-GPU parity evidence for the data-parallel product paths (NCCL over NVLink; run under torchrun on >= 2 GPUs
of one box: `gpurun --gpus 2 -- python -m torch.distributed.run --nproc-per-node 2 ... tools/dp_parity_check.py`):

  1. fused BatchNorm -> act-quant -> ReLU with GLOBAL-batch statistics (model/fused.py:_SyncBnActFn: the kernels' fp64
     (sum, sum of squares) all-reduced between the stats and apply launches, and (sum g_z, sum g_z xhat) in the
     backward) vs the single-device fused kernels on the gathered batch: differing codes, gx / ggamma / gbeta;
  2. the feature-sharded global-batch ADMM term (utils/dp_gram.py: all-to-all, partial Gram sums, one all-reduce,
     all-to-all back) vs the single-device fused quantizer + ADMM term on the gathered batch: y, D, trans_loss, gx;
  3. one data-parallel training step of resnet20_quant (QA) / resnet20_quant (QB + ADMM, feature mode) vs the
     single-device step on the gathered batch, beside the single-device step started 1e-7 away (the model's own
     sensitivity band: quantisation is discrete, rounding ties flip).

Writes gpurun_out/r02_dp_parity_n<world>.json on rank 0.  Evidence tool; CPU/gloo twins:
tests/test_dp_gram_gloo.py, tests/test_sharding_gloo.py; single-GPU composition twins: tests/test_gpu_dp.py,
tests/test_gpu_fused_bn.py::test_sync_bn_entries_compose_to_the_single_device_kernels."""
import copy
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alignq_b200 as aq  # noqa: E402
from alignq_b200.model.fused import bn_act  # noqa: E402
from alignq_b200.model.resnet import resnet20_quant  # noqa: E402
from alignq_b200.utils import dp_gram  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
dp_gram.configure()
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
out = {"world": world, "backend": "nccl"}
cl = lambda t_: t_.contiguous(memory_format=torch.channels_last)
rel = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def gather(t_):
    parts = [torch.empty_like(t_) for _ in range(world)]
    dist.all_gather(parts, t_.contiguous())
    return torch.cat(parts)


# ---- 1. SyncBN inside the fused kernels ---------------------------------------------------------------------------
for impl, shape, relu, res in [("nccl", (64, 16, 32, 32), True, False), ("nccl", (32, 64, 8, 8), True, True),
                               ("peer", (64, 16, 32, 32), True, False), ("peer", (32, 64, 8, 8), True, True),
                               ("peer", (16, 256, 4, 4), False, False)]:
    B, C, H, W = shape
    g = torch.Generator().manual_seed(100 + rank)
    x0 = cl((torch.randn(shape, generator=g) * 1.4 + 0.2).to(dev))
    gy = cl(torch.randn(shape, generator=g).to(dev))
    r0 = cl(torch.randn(shape, generator=g).to(dev)) if res else None
    torch.manual_seed(3)
    bn = torch.nn.BatchNorm2d(C).to(dev).train()
    with torch.no_grad():
        bn.weight.copy_(1 + 0.2 * torch.randn(C, device=dev))
        bn.bias.copy_(0.1 * torch.randn(C, device=dev))
    bn1 = copy.deepcopy(bn)
    q = aq.activation_quantize_fn(8, "second")
    aq.set_args(variant="A", act_range=2, abitW=8, fuse_bn_act=True, method="none", sync_bn=impl)
    for rep in range(3 if impl == "peer" else 1):           # peer mode: several exchanges through the 4-slot ring
        bn.weight.grad = None
        x = x0.clone().requires_grad_(True)
        rr = r0.clone().requires_grad_(True) if res else None
        if rep:
            bn.load_state_dict(bn1.state_dict())
        y = bn_act(bn, q, x, relu, residual=rr)
        (y * gy).sum().backward()
    gw = bn.weight.grad.clone()
    dist.all_reduce(gw)
    aq.set_args(sync_bn=False)
    xg, gyg = cl(gather(x0)), cl(gather(gy))
    rg = cl(gather(r0)).requires_grad_(True) if res else None
    xg = xg.requires_grad_(True)
    y1 = bn_act(bn1, q, xg, relu, residual=rg)
    (y1 * gyg).sum().backward()
    sl = slice(rank * B, (rank + 1) * B)
    bad = torch.tensor([float(((y.detach() - y1.detach()[sl]).abs() > 1e-6).sum())], device=dev)
    dist.all_reduce(bad)
    same = (y.detach() - y1.detach()[sl]).abs() <= 1e-6
    egx = torch.tensor([float(((x.grad - xg.grad[sl]).abs() * same).max() / xg.grad.abs().max())], device=dev)
    dist.all_reduce(egx, op=dist.ReduceOp.MAX)
    out[f"sync_bn_act[{impl}] {shape} relu={relu} res={res}"] = {
        "codes_differing": int(bad), "elements": B * world * C * H * W, "gx_max_err_over_max": float(egx),
        "ggamma_rel": rel(gw, bn1.weight.grad), "running_var_rel": rel(bn.running_var, bn1.running_var)}
    assert int(bad) <= max(2, int(1e-5 * B * world * C * H * W)) and float(egx) <= 1e-5 and rel(gw, bn1.weight.grad) <= 1e-4

# ---- 2. feature-sharded global-batch ADMM term -------------------------------------------------------------------
for Bg, shape, variant, mode in [(128, (16, 16, 16), "B", "tf32x3"), (128, (64, 8, 8), "B", "fp32"), (28 * world if 28 * world <= 128 else 112, (64, 14, 14), "C", "tf32x3")]:
    if Bg % world:
        continue
    b = Bg // world
    g = torch.Generator().manual_seed(200 + rank)
    x0 = torch.randn(b, *shape, generator=g).to(dev)
    gy = torch.randn(b, *shape, generator=g).to(dev)
    torch.manual_seed(5)
    admm = aq.ADMM(Bg).to(dev)
    admm1 = copy.deepcopy(admm)
    Fn = aq.activation_quantize_fn if variant == "B" else aq.activation_quantize_fn2
    aq.set_args(variant=variant, act_range=2, method="ours", gram_mode=mode, dp_gram="feature")
    x = x0.clone().requires_grad_(True)
    y, tl = Fn(8, "second", admm)(x)
    ((y * gy).sum() + 1.3 * tl).backward()
    aq.set_args(dp_gram="replica")
    xg, gyg = gather(x0).requires_grad_(True), gather(gy)
    y1, tl1 = Fn(8, "second", admm1)(xg)
    ((y1 * gyg).sum() + 1.3 * tl1).backward()
    sl = slice(rank * b, (rank + 1) * b)
    x64 = xg.detach().double().requires_grad_(True)
    from oracle import alignq_oracle as O                      # test tooling: fp64 autograd of the whole batch
    y64, l64, D64 = O.activation_quantize_admm(x64, 8, admm1.alterD.detach().double(), admm1.gamma.detach().double(), "second", variant, 2.0)
    ((y64 * gyg.double()).sum() + 1.3 * l64).backward()
    d = (x.grad.double() - x64.grad[sl]).abs()
    tol = 1e-5 * x64.grad[sl].abs() + 1e-6 * float(x64.grad.abs().max())
    worst = torch.tensor([float((d / tol).max())], device=dev)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    eps = 0.0 if variant == "B" else 1e-5
    gmax = float(O.corr(x64.detach().view(Bg, -1), x64.detach().view(Bg, -1), eps).abs().max())
    out[f"feature_sharded_admm B={Bg} {shape} {variant} {mode}"] = {
        "y_equal": bool(torch.equal(y.detach(), y1.detach()[sl])),
        "D_err_over_maxG_vs_single_device": float((admm.D - admm1.D).abs().max()) / gmax,
        "D_err_over_maxG_vs_fp64": float((admm.D.double() - D64).abs().max()) / gmax,
        "trans_loss_rel_err": abs(float(tl) - float(tl1)) / abs(float(tl1)),
        "gx_worst_over_1e-5_bar_vs_fp64_autograd": float(worst),
        "gx_rel_norm_vs_single_device": rel(x.grad, xg.grad[sl])}
    assert float(worst) <= 1.0 and float((admm.D.double() - D64).abs().max()) <= 1e-5 * gmax

# ---- 3. one training step, N ranks vs one device -------------------------------------------------------------------
from alignq_b200.utils.train import QATStep  # noqa: E402
for tag, variant, gram in (("resnet20_QA", "A", None), ("resnet20_QB_admm_feature", "B", "feature")):
    b = 32
    Bg = b * world
    g = torch.Generator().manual_seed(300 + rank)
    x = cl(torch.randn(b, 3, 32, 32, generator=g).to(dev))
    t = torch.randint(0, 10, (b,), generator=g).to(dev)
    xg, tg = cl(gather(x)), gather(t)
    res = {}
    for name in ("dp", "single", "single_perturbed"):
        aq.reset_args()
        aq.set_args(variant=variant, bitW=8, abitW=8, act_range=2, method="ours", gram_mode="tf32x3", fuse_bn_act=(variant == "A"),
                    train_batch_size=(Bg if gram else (b if name == "dp" else Bg)), sync_bn=("peer" if name == "dp" else False),
                    dp_gram=(gram if (gram and name == "dp") else "replica"))
        torch.manual_seed(0)
        m = resnet20_quant(8, 8, "second").to(dev).train()
        if name == "dp" and variant != "A":
            m = torch.nn.SyncBatchNorm.convert_sync_batchnorm(m)
        if name == "single_perturbed":
            with torch.no_grad():
                for p in m.parameters():
                    p.mul_(1.0 + 1e-7)
        st = QATStep(m, lr=0.04, momentum=0.9, weight_decay=1e-4, world_size=(world if name == "dp" else 1),
                     channels_last=True, single_backward=True)
        before = [p.detach().clone() for p in st.params]
        loss = st.step(x, t) if name == "dp" else st.step(xg, tg)
        loss = loss.detach().clone()
        if name == "dp":
            dist.all_reduce(loss)
            loss /= world
        upd = torch.cat([(p.detach() - q).reshape(-1) for p, q in zip(st.params, before)])
        res[name] = (float(loss), upd)
    out[f"train_step {tag}"] = {
        "loss_dp_mean": res["dp"][0], "loss_single": res["single"][0],
        "loss_rel_err": abs(res["dp"][0] - res["single"][0]) / abs(res["single"][0]),
        "loss_band_1e-7_perturbation": abs(res["single_perturbed"][0] - res["single"][0]) / abs(res["single"][0]),
        "update_rel_norm_err": rel(res["dp"][1], res["single"][1]),
        "update_band_1e-7_perturbation": rel(res["single_perturbed"][1], res["single"][1])}

if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/r02_dp_parity_n{world}.json", "w"), indent=1)
    print(json.dumps(out, indent=1))
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
