"""Benchmark of the AlignQ quantization hot path on B200 (contract: see the task prompt, "bench.py").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[0], the configuration the metric is quoted on): 8-bit ResNet-20
QAT, CIFAR-10 shaped synthetic data, per-GPU batch 128, variant QA (cdf_alignment/resnet-20-cifar-10),
SGD lr .04 mom .9 wd 1e-4 lam 1 lam2 4.  A "step" is one training iteration of the reference's
train() body: forward through the drop-in quantized modules, backward, SGD.step(idx, w_cdf, w_pdf,
lam, lam2).  N > 1: one process per GPU, batch sharded, one NCCL all-reduce of the flat gradient
buffer per step (weak scaling: per-GPU batch fixed).

One JSON line on stdout (rank 0):
  value     img/s, inputs resident in HBM, step replayed as a CUDA graph, CUDA-event time, max over ranks
  e2e       img/s through the same public API with HOST (pinned) inputs: H2D of the batch and a D2H
            read of the loss inside every timed step
  roofline  the CDF-quantizer fwd+bwd kernel pair on a 256 x 2^20 fp32 stream (4 GiB of traffic per
            pass, >> L2): algorithmic 20 B/elem / CUDA-event time, against MEASURED_PEAKS.json
  cpu_baseline  the oracle's restatement of the same training iteration on the host cores
--impl reference times only that CPU path (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "CDF-quant fwd+bwd HBM GB/s; 8-bit ResNet-20 QAT img/s @1/2/4/8 B200"
BATCH = 128
CONFIG = {"workload": "resnet20_quant W8A8 (QA) CIFAR-10 synthetic 32x32, QAT step fwd+bwd+SGD.step",
          "per_gpu_batch": BATCH, "variant": "A", "bitW": 8, "abitW": 8, "act_range": 2,
          "optimizer": "SGD lr=0.04 momentum=0.9 wd=1e-4 lam=1 lam2=4"}
FALLBACK_HBM_GBS = 6650.0


def peaks():
    try:
        p = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic_bytes():
    """DRAM bytes of one fwd + one bwd launch from the committed `ncu --set full` capture (or None)."""
    try:
        rows = json.load(open(os.path.join(REPO, "profiles", "r02_ncu_full_act_kernels.json")))
        tot = 0.0
        for r in rows:
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                v, unit = r[k].split()
                tot += float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit]
        return tot
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle's port of the reference training iteration (kind = "port")
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, budget_s=150.0, batch=BATCH, workload="resnet20"):
    """The workload's batch is NEVER changed: when the box is slow the number of timed iterations
    is cut instead (never below 10, warm-up never below 2), and the sample string says how many ran."""
    import torch
    from oracle import models_oracle as MO

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    if workload == "resnet56_admm":                        # configs[1]: QB + ADMM (oracle pinned by make_model_golden.py)
        model = MO.OracleResNet([9, 9, 9], 8, 8, "B", 2.0, dim=batch).train()
    elif workload == "densenet40":                         # configs[3] (oracle pinned by make_model_golden.py --job densenet40_A)
        model = MO.OracleDenseNet(8, 8, 2.0).train()
    elif workload == "mobilenetv2":                        # configs[2]: W4A4, batch 256 (pinned by --job mobilenetv2_A)
        model = MO.OracleMobileNetV2(4, 4, 2.0).train()
    else:
        model = MO.resnet20_oracle(8, 8, "A", act_range=2.0, dim=batch).train()
    bits = 4 if workload == "mobilenetv2" else 8
    tr = MO.OracleTrainer(model, lr=0.04, momentum=0.9, weight_decay=1e-4, lam=1.0, lam2=4.0, bitW=bits)
    x = torch.randn(batch, 3, 32, 32)
    t = torch.randint(0, 10, (batch,))
    t0 = time.perf_counter()
    tr.step(x, t)                                          # first (untimed) iteration sizes the sample
    first = time.perf_counter() - t0
    warm = max(warmup - 1, 1)
    if first * (steps + warm) > budget_s:                  # bound the run: fewer iterations, same batch
        warm = min(warm, 2)
        steps = max(10, min(steps, int(budget_s / first) - warm))
    for _ in range(warm):
        tr.step(x, t)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.step(x, t)
    dt = time.perf_counter() - t0
    return {"value": batch * steps / dt, "unit": "img/s", "cores": cores, "kind": "port",
            "sample": f"{steps} timed iterations of the workload's batch {batch} after {warm + 1} warm-up, "
                      f"oracle/models_oracle.OracleTrainer ({workload}) on CPU, torch threads={cores}",
            "ms_per_step": dt / steps * 1e3, "steps_timed": steps}


def gpu_eager_port_run(dev, steps=10, warmup=3, batch=BATCH):
    """The same oracle restatement of the reference iteration, run in GPU EAGER mode (the ATen/cuDNN kernels the
    reference itself launches on a GPU): the "same box, reference's own GPU path" bar beside the CPU baseline.
    A reported baseline only -- nothing of the product goes through it."""
    import torch
    from oracle import models_oracle as MO

    torch.manual_seed(0)
    model = MO.resnet20_oracle(8, 8, "A", act_range=2.0, dim=batch).to(dev).train()
    tr = MO.OracleTrainer(model, lr=0.04, momentum=0.9, weight_decay=1e-4, lam=1.0, lam2=4.0, bitW=8)
    x = torch.randn(batch, 3, 32, 32, device=dev)
    t = torch.randint(0, 10, (batch,), device=dev)
    for _ in range(warmup):
        tr.step(x, t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        tr.step(x, t)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": batch / (ms * 1e-3), "unit": "img/s", "ms_per_step": ms,
            "sample": f"{steps} iterations of batch {batch} after {warmup} warm-up, oracle/models_oracle.OracleTrainer "
                      "with device=cuda (PyTorch eager, NCHW, cuDNN), CUDA events"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload if args.workload in ("resnet20", "resnet56_admm", "densenet40", "mobilenetv2") else "resnet20"
    ref_batch = 256 if wl == "mobilenetv2" else BATCH
    cb = cpu_reference_run(args.steps, args.warmup, batch=ref_batch, workload=wl)
    # same config keys as the product arm's line (the CPU arm has one process whatever N is: it does not scale)
    cfg = dict(CONFIG, global_batch=ref_batch, parallelism="cpu", cpu_steps_timed=cb["steps_timed"])
    if wl == "mobilenetv2":
        cfg.update(workload="mobile_v2 W4A4 (QA) SVHN synthetic 32x32, QAT step", bitW=4, abitW=4, per_gpu_batch=ref_batch)
    elif wl == "resnet56_admm":
        cfg.update(workload="resnet56_quant W8A8 (QB) + ADMM, CIFAR-10 synthetic, QAT step", variant="B")
    elif wl == "densenet40":
        cfg.update(workload="densenet_40_quant W8A8 (QA) CIFAR-10 synthetic, QAT step")
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "img/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def dp_parity_block(args, world, rank, dev, make_model, batch, img_hw, ncls, fuse):
    """N ranks vs ONE device on the same global batch: every rank runs forward + backward on its shard
    with the global-batch modes on (fused-kernel SyncBN; dp_gram as requested), gradients are summed over ranks and
    divided by N; then every rank runs the SAME model on the gathered global batch alone and compares.  Relative errors
    of the loss, trans_loss, the first ADMM layer's D and the flat parameter gradient.  (Quantisation is discrete: a BN
    output on a rounding tie may flip a code between the two runs; see DESIGN.md 2 for the bands.)"""
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F_
    import alignq_b200 as aq
    from alignq_b200.utils.admm import ADMM
    fmt = torch.contiguous_format if args.nchw else torch.channels_last
    g = torch.Generator().manual_seed(777 + rank)
    x = torch.randn(batch, 3, img_hw, img_hw, generator=g).to(dev).contiguous(memory_format=fmt)
    t = torch.randint(0, ncls, (batch,), generator=g).to(dev)
    xs = [torch.empty_like(x) for _ in range(world)]
    ts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(xs, x)
    dist.all_gather(ts, t)
    xg, tg = torch.cat(xs).contiguous(memory_format=fmt), torch.cat(ts)
    saved = {k: getattr(aq.args, k) for k in ("sync_bn", "dp_gram", "train_batch_size", "fuse_bn_act")}
    # heuristic (not auto-tuned) fp32 convolutions for both passes: with cudnn.benchmark the two batch sizes may get
    # different algorithms, whose last-bit differences the 21 quantized layers amplify (the 1e-7 band below)
    saved_backend = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = False, False, False
    feature = args.dp_gram == "feature"
    out = {}

    def run(model, xx, tt, tl_scale):
        for p in model.parameters():
            p.grad = None
        o = model(xx)
        logits, tl = o if isinstance(o, tuple) else (o, None)
        ce = F_.cross_entropy(logits, tt)
        total = ce if not torch.is_tensor(tl) else ce + tl * tl_scale
        total.backward()
        flat = torch.cat([p.grad.reshape(-1) for n, p in model.named_parameters()
                          if p.grad is not None and "alterD" not in n and "gamma" not in n])
        D = next((m.D.clone() for m in model.modules() if isinstance(m, ADMM) and getattr(m, "D", None) is not None), None)
        return ce.detach(), (tl.detach() if torch.is_tensor(tl) else None), D, flat

    # N-rank pass
    aq.set_args(sync_bn=args.sync_bn_impl, dp_gram=args.dp_gram, fuse_bn_act=fuse, train_batch_size=(batch * world if feature else batch))
    torch.manual_seed(0)
    m_dp = make_model().to(dev).train()
    if not fuse_effective(args, fuse):
        m_dp = torch.nn.SyncBatchNorm.convert_sync_batchnorm(m_dp)
    if fmt is torch.channels_last:
        m_dp = m_dp.to(memory_format=torch.channels_last)
    ce, tl, D, flat = run(m_dp, x, t, float(world) if feature else 1.0)
    dist.all_reduce(flat)
    flat /= world
    cem = ce.clone()
    dist.all_reduce(cem)
    cem /= world
    # single-device pass on the gathered global batch (only meaningful for the global-batch modes)
    aq.set_args(sync_bn=False, dp_gram="replica", train_batch_size=(batch * world if feature else batch))
    torch.manual_seed(0)
    m_sd = make_model().to(dev).train()
    if fmt is torch.channels_last:
        m_sd = m_sd.to(memory_format=torch.channels_last)
    has_admm = any(isinstance(m, ADMM) for m in m_sd.modules())
    if has_admm and not feature:                            # replica mode: the per-rank [b,b] Gram is NOT the global one
        out["note"] = "dp_gram=replica: each rank = the reference at batch b; global-batch parity needs --dp-gram feature"
    else:
        ce1, tl1, D1, flat1 = run(m_sd, xg, tg, 1.0)
        rel = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
        out.update(ce_rel_err=abs(float(cem) - float(ce1)) / abs(float(ce1)), grad_rel_norm_err=rel(flat, flat1))
        with torch.no_grad():                               # the model's own sensitivity: the same single-device pass started
            for p in m_sd.parameters():                     # 1e-7 away (rounding ties of the 21 quantized layers flip)
                p.mul_(1.0 + 1e-7)
        ce2, _, _, flat2 = run(m_sd, xg, tg, 1.0)
        out.update(band_ce_rel_1e7_perturbation=abs(float(ce2) - float(ce1)) / abs(float(ce1)),
                   band_grad_rel_norm_1e7_perturbation=rel(flat2, flat1))
        if tl1 is not None:
            out.update(trans_loss_rel_err=abs(float(tl) - float(tl1)) / abs(float(tl1)),
                       D_first_layer_err_over_max=float((D - D1).abs().max() / D1.abs().max()))
    out.update(world=world, per_gpu_batch=batch, global_batch=batch * world,
               what="N-rank step (global-batch BN statistics%s) vs the same model on the gathered global batch on one GPU"
                    % (", feature-sharded global Gram" if feature else ""))
    aq.set_args(**saved)
    torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved_backend
    del m_dp, m_sd
    torch.cuda.empty_cache()
    return out


def fuse_effective(args, fuse):
    """True when every BatchNorm -> act-quant pair of the workload runs in the fused kernels (which carry their own
    global-batch statistics exchange); the ADMM workloads and --nchw / --no-fuse use separate modules."""
    return fuse and args.workload in ("resnet20", "mobilenetv2", "densenet40")


def fused_code_mismatch(dev):
    """Fraction of activation codes of the fused BatchNorm -> act-quant -> ReLU kernels that differ (by one level: a BN
    output on a rounding tie) from cuDNN BatchNorm + the stand-alone quantizer kernel, on the three layer shapes of the
    headline workload; beside it the same count for cuDNN's fp32 BN against an fp64 BN (the reference's own tie floor)."""
    import copy
    import torch
    import alignq_b200 as aq
    from alignq_b200.model.fused import bn_act
    out = {}
    worst = 0.0
    for shape in ((BATCH, 16, 32, 32), (BATCH, 32, 16, 16), (BATCH, 64, 8, 8)):
        torch.manual_seed(0)
        x = (torch.randn(shape, device=dev) * 1.5 + 0.3).contiguous(memory_format=torch.channels_last)
        bn = torch.nn.BatchNorm2d(shape[1]).to(dev).train()
        bn32, bn64 = copy.deepcopy(bn), copy.deepcopy(bn).double()
        q = aq.activation_quantize_fn(8, "second")
        with torch.no_grad():
            y = bn_act(bn, q, x, True)
            aq.set_args(fuse_bn_act=False)
            y32 = torch.relu(q(bn32(x)))
            z64 = bn64(x.double())
            y64 = torch.relu(q(z64.float()))               # fp64 BN rounded once to fp32, then the same quantizer kernel
            aq.set_args(fuse_bn_act=True)
        n = y.numel()
        f = float(((y - y32).abs() > 1e-6).sum()) / n
        worst = max(worst, f)
        out["x".join(map(str, shape))] = {"vs_cudnn_bn": f, "vs_fp64_bn": float(((y - y64).abs() > 1e-6).sum()) / n,
                                          "cudnn_bn_vs_fp64_bn": float(((y32 - y64).abs() > 1e-6).sum()) / n}
    return {"code_mismatch_frac_max": worst, "bar": 1e-5, "per_shape": out,
            "what": "+-1 code at BatchNorm-output rounding ties, fused kernels vs cuDNN BN + quantizer kernel"}


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for ln in self.f:
            c = [v.strip() for v in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def run_product(args):
    import torch
    import torch.distributed as dist

    import alignq_b200 as aq
    from alignq_b200 import _lib as L
    from alignq_b200.model.resnet import resnet20_quant
    from alignq_b200.utils.train import HostFeeder, QATStep

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = aq.load_library()
    torch.backends.cudnn.benchmark = os.environ.get("ALIGNQ_CUDNN_BENCHMARK", "1") != "0"   # 0: cheap ncu runs
    fuse = not (args.no_fuse or args.nchw)
    aq.set_args(variant="A", bitW=8, abitW=8, act_range=2, lam=1.0, lam2=4.0, method="ours", train_batch_size=BATCH,
                fuse_bn_act=fuse, own_conv=("off" if args.nchw else args.own_conv), fused_head=not (args.no_fused_head or args.nchw),
                own_conv_channels=tuple(int(c) for c in args.own_conv_channels.split(",") if c),
                own_wgrad_channels=tuple(int(c) for c in args.own_wgrad_channels.split(",") if c),
                own_dgrad_channels=tuple(int(c) for c in args.own_dgrad_channels.split(",") if c))
    torch.manual_seed(0)                                   # identical replicas on every rank
    batch, img_hw, ncls, forward_loss = BATCH, 32, 10, None
    # ADMM(dim): the GLOBAL batch in dp_gram='feature' mode (weak scaling: per-GPU batch x ranks), else the per-GPU batch
    admm_dim = lambda b: b * world if (args.dp_gram == "feature" and world > 1 and not args.strong) else b
    make_model = None
    if args.workload == "resnet20":
        make_model = lambda: resnet20_quant(8, 8, "second")
        model = make_model()
    elif args.workload == "resnet56_admm":                 # configs[1]: QB, W8A8 + ADMM correlation preservation
        from alignq_b200.model.resnet import resnet56_quant
        aq.set_args(variant="B", gram_mode=args.gram_mode, fuse_bn_act=False, train_batch_size=admm_dim(BATCH))
        make_model = lambda: resnet56_quant(8, 8, "second")
        model = make_model()
        CONFIG.update(workload=f"resnet56_quant W8A8 (QB) + ADMM, gram_mode={args.gram_mode}, CIFAR-10 synthetic, QAT step", variant="B")
    elif args.workload == "mobilenetv2":                   # configs[2]: W4A4, depthwise convs, batch 256
        from alignq_b200.model.mobilenetV2 import mobile_v2
        batch = 256
        aq.set_args(bitW=4, abitW=4, train_batch_size=batch)
        model = mobile_v2(4, 4, "second")
        CONFIG.update(workload="mobile_v2 W4A4 (QA) SVHN synthetic 32x32, QAT step", bitW=4, abitW=4, per_gpu_batch=batch)
    elif args.workload == "resnet50_dann":                 # configs[4]: QC, Office-31 shaped 224x224, batch 28 per GPU,
        from alignq_b200.model.dann import resnet50_dann   # source + target forward per iteration (main.py:372-385)
        batch = 28
        aq.set_args(variant="C", gram_mode=args.gram_mode, fuse_bn_act=False, train_batch_size=admm_dim(batch))
        model = resnet50_dann(8, 8, "second")
        CONFIG.update(workload=f"resnet50_dann W8A8 (QC) + ADMM, gram_mode={args.gram_mode}, Office-31 synthetic 224x224, "
                               "source+target forward, QAT step", variant="C", per_gpu_batch=batch)
        img_hw, ncls = 224, 31

        def forward_loss(m, x, t):
            b = x.shape[0] // 2
            cls_s, dom_s, tl_s = m(x[:b], 0.5)
            _, dom_t, tl_t = m(x[b:], 0.5)
            zeros, ones = torch.zeros_like(t), torch.ones_like(t)
            F_ = torch.nn.functional
            return F_.cross_entropy(cls_s, t) + F_.cross_entropy(dom_s, zeros) + F_.cross_entropy(dom_t, ones), tl_s + tl_t
    else:                                                  # configs[3]: DenseNet-40 (k=12) W8A8
        from alignq_b200.model.densenet import densenet_40_quant
        model = densenet_40_quant(8, 8, "second")
        CONFIG.update(workload="densenet_40_quant W8A8 (QA) CIFAR-10 synthetic, QAT step")
    if args.strong:
        if batch % world:
            raise SystemExit(f"--strong: global batch {batch} is not divisible by {world} ranks")
        batch //= world
        if not (args.dp_gram == "feature" and world > 1):  # feature mode keeps ADMM(dim = global batch)
            aq.set_args(train_batch_size=batch)
        CONFIG.update(per_gpu_batch=batch)
    model = model.to(dev).train()
    sync_bn = world > 1 and not args.local_bn
    if world > 1:
        from alignq_b200.utils import dp_gram
        dp_gram.configure()                                # process group of the global-batch modes
        aq.set_args(sync_bn=(args.sync_bn_impl if sync_bn else False), dp_gram=args.dp_gram)
        if args.dp_gram == "feature":                      # ADMM dim = GLOBAL batch (the model was built with it below)
            CONFIG.update(dp_gram="feature: global-batch Gram (all-to-all + partial sums + one all-reduce per layer)")
    if sync_bn and not fuse_effective(args, fuse):
        # layers the fused bn-act kernels do not cover (ADMM variants, NCHW): torch's SyncBatchNorm
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    step = QATStep(model, lr=0.04, momentum=0.9, weight_decay=1e-4, world_size=world, channels_last=not args.nchw,
                   single_backward=True, forward_loss=forward_loss)

    g = torch.Generator().manual_seed(1234 + rank)         # each rank its own shard of the synthetic batch
    n_host = 8
    nimg = batch * (2 if forward_loss is not None else 1)      # DANN: source + target images per step
    n_host = 8 if img_hw == 32 else 2
    host_x = [torch.randn(nimg, 3, img_hw, img_hw, generator=g).pin_memory() for _ in range(n_host)]
    host_t = [torch.randint(0, ncls, (batch,), generator=g).pin_memory() for _ in range(n_host)]
    fmt = torch.contiguous_format if args.nchw else torch.channels_last
    host_x = [h.contiguous(memory_format=fmt).pin_memory() for h in host_x]
    dev_x = [h.to(dev) for h in host_x]
    dev_t = [h.to(dev) for h in host_t]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    if args.kernel_shares:                                 # dev aid: CUPTI kernel-time shares of two eager steps, then exit
        import collections
        from torch.profiler import profile, ProfilerActivity
        for _ in range(4):
            step.step(dev_x[0], dev_t[0])
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(2):
                step.step(dev_x[0], dev_t[0])
            torch.cuda.synchronize()
        agg = collections.defaultdict(lambda: [0, 0.0])
        for e in prof.events():
            if getattr(e.device_type, "name", "") == "CUDA":
                agg[e.name[:96]][0] += 1
                agg[e.name[:96]][1] += e.device_time
        tot = sum(v[1] for v in agg.values())
        lines = [f"# {args.workload}: kernel time per step {tot / 2:.1f} us over {sum(v[0] for v in agg.values()) // 2} kernels (eager, torch.profiler)"]
        lines += [f"{d / 2:9.1f} us {100 * d / tot:5.1f}% x{c // 2:4d}  {k}" for k, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]]
        os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
        open(os.path.join(REPO, "gpurun_out", f"kernel_shares_{args.workload}.txt"), "w").write("\n".join(lines) + "\n")
        print("\n".join(lines), file=sys.stderr)
        os._exit(0)

    graphed = not args.no_graph
    launches_per_step = None
    if graphed:
        try:
            c0 = lib.alignq_launch_count()
            step.capture(dev_x[0], dev_t[0], warmup=3)
            # 3 eager warm-up iterations + the captured one
            launches_per_step = (lib.alignq_launch_count() - c0) // 4
        except Exception as e:                              # pragma: no cover - reported, not hidden
            if rank == 0:
                print(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); running eager", file=sys.stderr)
            graphed = False
            step.graph = None
    if launches_per_step is None:
        c0 = lib.alignq_launch_count()
        step.step(dev_x[0], dev_t[0])
        launches_per_step = lib.alignq_launch_count() - c0

    if args.timeline:                                      # dev aid: CUPTI timeline of graph replays (real overlap, real gaps), then exit
        import collections
        from torch.profiler import profile, ProfilerActivity
        for _ in range(5):
            flush.zero_()
            step.step(dev_x[0], dev_t[0])
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                flush.zero_()
                step.step(dev_x[0], dev_t[0])
                torch.cuda.synchronize()
        evs = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events()
                      if getattr(e.device_type, "name", "") == "CUDA"), key=lambda r: r[0])
        # split at the L2-flush fills (the only 256 MiB FillFunctor<unsigned char> launches)
        steps_, cur = [], []
        for a, b, n in evs:
            if "FillFunctor<unsigned char>" in n:
                if cur:
                    steps_.append(cur)
                cur = []
            else:
                cur.append((a, b, n))
        if cur:
            steps_.append(cur)
        last = steps_[-1]
        t0, t1 = last[0][0], max(b for _, b, _ in last)
        busy, cover_end, agg = 0.0, t0, collections.defaultdict(lambda: [0, 0.0])
        for a, b, n in last:
            if b > cover_end:
                busy += b - max(a, cover_end)
                cover_end = b
            agg[n[:90]][0] += 1
            agg[n[:90]][1] += b - a
        lines = [f"# {args.workload}: one graph replay under CUPTI: span {t1 - t0:.1f} us, device busy (union over streams) {busy:.1f} us, "
                 f"idle {t1 - t0 - busy:.1f} us, sum of kernel durations {sum(v[1] for v in agg.values()):.1f} us, {len(last)} kernels"]
        lines += [f"{d:9.1f} us x{c:4d}  {k}" for k, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]]
        lines.append("# timeline (start us, dur us, name)")
        lines += [f"{a - t0:9.1f} {b - a:7.1f}  {n[:100]}" for a, b, n in last]
        os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
        open(os.path.join(REPO, "gpurun_out", f"timeline_{args.workload}.txt"), "w").write("\n".join(lines) + "\n")
        print("\n".join(lines[:50]), file=sys.stderr)
        os._exit(0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(k, w, host_inputs, st=None):
        """k steps after w warm-ups; per-step CUDA events on the launching stream, L2 flushed between
        steps (outside the events).  Returns summed device milliseconds.
        host_inputs: every step's batch starts in pinned HOST memory and is copied to the device inside the timed
        region -- through HostFeeder, i.e. the copy of batch i+1 is issued (copy stream) right before step i is
        launched and overlaps it; step i waits for ITS copy.  In steady state every timed step contains exactly one
        H2D of a full batch and one D2H read of the loss."""
        st = step if st is None else st
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        sink = 0.0
        feeder = None
        if host_inputs:
            feeder = HostFeeder(st, dev_x[0], dev_t[0])
            feeder.prefetch(host_x[0], host_t[0])
        for i in range(w + k):
            flush.zero_()
            j = i % n_host
            if i >= w:
                ev[i - w][0].record()
            if host_inputs:
                jn = (i + 1) % n_host
                loss = feeder.step(host_x[jn], host_t[jn])
                sink += float(loss.item())                  # D2H read of the step's loss (4 bytes), syncs
            else:
                st.step(dev_x[j], dev_t[j])
            if i >= w:
                ev[i - w][1].record()
            if i == w - 1:
                barrier()
        barrier()
        return sum(a.elapsed_time(b) for a, b in ev), sink

    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ms, _ = timed(args.steps, args.warmup, host_inputs=False)
    clocks = sampler.stop() if sampler else None
    ms_e2e, _ = timed(args.steps, max(args.warmup, 3), host_inputs=True)
    tt = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(tt[0]), float(tt[1])
    value = world * batch * args.steps / (ms * 1e-3)
    e2e = world * batch * args.steps / (ms_e2e * 1e-3)

    roofline = cpu = gram = None
    if rank == 0 and not args.lean:
        # ---- roofline leg: the CDF-quantizer kernel pair on a stream far larger than L2 -----------
        n = 256 * (1 << 20)
        x = torch.randn(n, device=dev)
        gy = torch.randn(n, device=dev)
        y, gx = torch.empty_like(x), torch.empty_like(x)
        s = L.stream_ptr()
        reps, tf, tb = 10, 0.0, 0.0
        for i in range(3 + reps):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            L.check(lib.alignq_act_fwd(x.data_ptr(), y.data_ptr(), 0, n, 8, 2.0, 0, 0, s), "act_fwd")
            e[1].record()
            L.check(lib.alignq_act_bwd(x.data_ptr(), gy.data_ptr(), gx.data_ptr(), n, 8, 2.0, 0, 0, s), "act_bwd")
            e[2].record()
            torch.cuda.synchronize()
            if i >= 3:
                tf += e[0].elapsed_time(e[1]) / reps
                tb += e[1].elapsed_time(e[2]) / reps
        peak, how = peaks()
        achieved = 20.0 * n / ((tf + tb) * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": ncu_traffic_bytes(), "traffic_source": "profiles/r02_ncu_full_act_kernels.json "
                    "(dram__bytes_read.sum + dram__bytes_write.sum of one fwd + one bwd launch on this input; "
                    "algorithmic = 5.369e9 B)", "peak_source": how,
                    "kernel": "act_fwd_vec_kernel + act_bwd_vec_kernel (CDF quantizer fwd + fused STE bwd)",
                    "algorithmic_bytes_per_elem": 20, "elems_per_launch": n,
                    "fwd": {"ms": tf, "gbs": 8.0 * n / (tf * 1e-3) / 1e9, "frac": 8.0 * n / (tf * 1e-3) / 1e9 / peak},
                    "bwd": {"ms": tb, "gbs": 12.0 * n / (tb * 1e-3) / 1e9, "frac": 12.0 * n / (tb * 1e-3) / 1e9 / peak},
                    "input": "256 x 2^20 fp32 (1 GiB per tensor, inputs >> L2), W8A8 variant A"}
        del x, gy, y, gx
        # ---- tensor leg: the ADMM Gram as a bf16 SYRK on tcgen05 (SURVEY 8d micro-shape B=256, F=2^20) ----
        try:
            Bg, Fg = 256, 1 << 20
            xg = torch.randn(Bg, Fg, device=dev).to(torch.bfloat16)
            Gg = torch.empty(Bg, Bg, device=dev)
            wsg = torch.empty(int(lib.alignq_gram_bf16_ws_bytes(Bg)), dtype=torch.uint8, device=dev)
            call = lambda: L.check(lib.alignq_gram_bf16(xg.data_ptr(), Bg, Fg, 1, Gg.data_ptr(), wsg.data_ptr(), wsg.numel(), s), "gram_bf16")
            for _ in range(3):
                call()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):                           # back to back: the 537 MB input cannot stay in the 126 MB L2
                call()
            e1.record()
            torch.cuda.synchronize()
            tg = e0.elapsed_time(e1) / reps
            pk = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else {}
            tf = 2.0 * Bg * Bg * Fg / (tg * 1e-3) / 1e12
            gram = {"bound": "tensor", "achieved": tf, "unit": "TFLOP/s", "peak": pk.get("bf16_tflops", 1590.0),
                    "frac": tf / pk.get("bf16_tflops", 1590.0), "frac_of_sustained": tf / pk.get("bf16_tflops_sustained", 1400.0),
                    "ms": tg, "kernel": "gram_bf16_ldgsts_kernel (cp.async producers with hand-applied 128B swizzle + tcgen05.mma kind::f16, M=128 N=256/128, split-K) + reduce",
                    "shape": [Bg, Fg], "flops": "2*B^2*F (full product)", "tolerance": "1e-2 rel vs fp32 inputs (bf16 operands)",
                    "input": "bf16 [256, 2^20] (537 MB >> L2)"}
            del xg, Gg, wsg
        except Exception as e:                              # pragma: no cover
            gram = {"error": str(e)[:200]}
        if world == 1 and not args.no_cpu_baseline and args.workload in ("resnet20", "resnet56_admm", "densenet40", "mobilenetv2"):
            cb = cpu_reference_run(20, 3, budget_s=40.0, batch=batch, workload=args.workload)
            cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            if args.workload == "resnet20":
                try:
                    cpu["gpu_eager_port"] = gpu_eager_port_run(dev)
                except Exception as e:                      # pragma: no cover - a baseline, never fatal
                    cpu["gpu_eager_port"] = {"error": str(e)[:200]}

    dp_parity = None
    if world > 1 and make_model is not None and not args.no_dp_parity and (sync_bn or args.dp_gram == "feature"):
        try:
            dp_parity = dp_parity_block(args, world, rank, dev, make_model, batch, img_hw, ncls, fuse)
        except Exception as e:                              # pragma: no cover - evidence block, never fatal
            dp_parity = {"error": f"{type(e).__name__}: {str(e)[:300]}"}

    local_bn = None
    if world > 1 and sync_bn and args.workload == "resnet20" and not args.no_local_bn_line:
        # the same data-parallel step with PER-RANK BatchNorm statistics (plain DDP semantics: every rank = the reference
        # at batch 128) and the gradient all-reduce overlapped with the backward -- beside the global-batch number above
        aq.set_args(sync_bn=False)
        torch.manual_seed(0)
        m3 = make_model().to(dev).train()
        st3 = QATStep(m3, lr=0.04, momentum=0.9, weight_decay=1e-4, world_size=world, channels_last=not args.nchw,
                      single_backward=True)
        if graphed:
            st3.capture(dev_x[0], dev_t[0], warmup=3)
        k3 = min(args.steps, 100)
        ms3, _ = timed(k3, 5, host_inputs=False, st=st3)
        t3 = torch.tensor([ms3], dtype=torch.float64, device=dev)
        dist.all_reduce(t3, op=dist.ReduceOp.MAX)
        local_bn = {"value": world * batch * k3 / (float(t3[0]) * 1e-3), "unit": "img/s", "ms_per_step": float(t3[0]) / k3, "steps": k3,
                    "what": "same N-GPU step with per-rank BatchNorm statistics (no statistics exchange) and the gradient "
                            "all-reduce overlapped with the backward on the side stream"}
        aq.set_args(sync_bn=args.sync_bn_impl)

    fused_parity = no_fuse = None
    if rank == 0 and world == 1 and fuse and args.workload == "resnet20" and not args.lean:
        fused_parity = fused_code_mismatch(dev)
        # the same step with BatchNorm / quantizer / ReLU as separate kernels, beside the fused number
        aq.set_args(fuse_bn_act=False)
        torch.manual_seed(0)
        m2 = resnet20_quant(8, 8, "second").to(dev).train()
        st2 = QATStep(m2, lr=0.04, momentum=0.9, weight_decay=1e-4, channels_last=not args.nchw, single_backward=True)
        if graphed:
            st2.capture(dev_x[0], dev_t[0], warmup=3)
        k2 = min(args.steps, 50)
        ms2, _ = timed(k2, 5, host_inputs=False, st=st2)
        no_fuse = {"value": batch * k2 / (ms2 * 1e-3), "unit": "img/s", "ms_per_step": ms2 / k2, "steps": k2,
                   "what": "same workload with fuse_bn_act=False (cuDNN BatchNorm + act-quant kernel + ReLU)"}
        aq.set_args(fuse_bn_act=True)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong" if args.strong else "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(CONFIG, global_batch=batch * world, parallelism=f"dp{world}", cuda_graph=graphed,
                               activation_layout="nchw" if args.nchw else "channels_last", fused_bn_act=fuse,
                               conv3x3=("cuDNN" if (args.nchw or args.own_conv == "off") else
                                        f"own tcgen05 kernels ({args.own_conv}) for 3x3/s1/Cin==Cout in ({args.own_conv_channels}) [+ data gradient in ({args.own_dgrad_channels}), "
                                        f"weight gradient in ({args.own_wgrad_channels})] and own fp32 first-layer kernels; cuDNN (tf32) for the rest"),
                               fused_head=not (args.no_fused_head or args.nchw),
                               sync_bn=bool(sync_bn),
                               sync_bn_impl=(None if not sync_bn else (
                                   ("fused bn-act kernels, fp64 (sum, sumsq) exchanged inside the kernels over NVLink peer memory"
                                    if args.sync_bn_impl == "peer" else "fused bn-act kernels, fp64 (sum, sumsq) NCCL all-reduce")
                                   if fuse_effective(args, fuse) else "torch.nn.SyncBatchNorm")),
                               l2="flushed between timed steps (256 MiB memset, outside the event pairs)"),
                "e2e": {"value": e2e, "unit": "img/s", "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": nimg * 3 * img_hw * img_hw * 4 + batch * 8, "d2h_bytes_per_step": 4},
                "gpu_launches": int(launches_per_step) * args.steps,
                "gpu_launches_per_step": int(launches_per_step),
                "clocks": clocks, "roofline": roofline, "gram_tensor_roofline": gram, "cpu_baseline": cpu,
                "fused_bn_act_parity": fused_parity, "no_fuse": no_fuse, "dp_parity": dp_parity, "local_bn": local_bn}
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear down without NCCL's communicator destructor: with captured NCCL kernels still referenced
        # by the CUDA graph, destroy_process_group() was seen to hang on the box.  Results are out.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", type=str, default="resnet20", choices=["resnet20", "resnet56_admm", "mobilenetv2", "densenet40", "resnet50_dann"],
                    help="default resnet20 = BASELINE.json configs[0] (the metric's workload); the others are configs[1..3], "
                    "for the record only (their JSON line says so in config.workload)")
    ap.add_argument("--gram-mode", type=str, default="tf32x3", choices=["fp32", "tf32x3", "bf16"])
    ap.add_argument("--no-fused-head", action="store_true", help="library kernels for avg-pool -> classifier -> cross-entropy")
    ap.add_argument("--timeline", action="store_true", help="CUPTI timeline of one graph replay -> gpurun_out/timeline_<workload>.txt, exit")
    ap.add_argument("--kernel-shares", action="store_true", help="profile two eager steps, write gpurun_out/kernel_shares_<workload>.txt, exit")
    ap.add_argument("--strong", action="store_true", help="strong scaling: the GLOBAL batch stays 128 (per-GPU batch 128/N); "
                    "default is weak scaling (per-GPU batch 128)")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--sync-bn", action="store_true", help="(default for N>1) global-batch BatchNorm statistics")
    ap.add_argument("--sync-bn-impl", type=str, default="peer", choices=["peer", "nccl"],
                    help="how the fused bn-act kernels exchange their global-batch sums: inside the kernels over NVLink peer "
                    "memory (default) or with an NCCL all-reduce between the launches")
    ap.add_argument("--local-bn", action="store_true", help="N>1: per-rank BatchNorm statistics (plain DDP semantics) "
                    "instead of the default global-batch statistics")
    ap.add_argument("--dp-gram", type=str, default="replica", choices=["replica", "feature"],
                    help="N>1, ADMM workloads: per-rank [b,b] Gram (reference at batch b) or the global-batch Gram")
    ap.add_argument("--own-conv", type=str, default="tf32", choices=["off", "tf32", "tf32x3"],
                    help="3x3 / stride-1 quantized convolutions on the hand-written tcgen05 kernels: tf32 (one pass, the numerics "
                    "of cuDNN under torch's default allow_tf32), tf32x3 (fp32 parity) or off (cuDNN everywhere)")
    ap.add_argument("--own-conv-channels", type=str, default="16", help="channel counts (Cin == Cout) routed to the own kernels")
    ap.add_argument("--own-dgrad-channels", type=str, default="32", help="further channel counts whose DATA gradient alone runs on the own kernel")
    ap.add_argument("--own-wgrad-channels", type=str, default="32", help="further channel counts whose WEIGHT gradient alone runs on the own kernel")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lean", action="store_true", help="profiler runs: the timed step only (no roofline / Gram / no-fuse / parity legs)")
    ap.add_argument("--no-local-bn-line", action="store_true", help="N>1: skip the extra timing with per-rank BatchNorm statistics")
    ap.add_argument("--no-dp-parity", action="store_true", help="N>1: skip the N-rank vs single-device parity block")
    ap.add_argument("--no-fuse", action="store_true", help="run BatchNorm / quantizer / ReLU as separate kernels "
                    "(default with channels_last: the fused bn->act-quant->relu kernels, SURVEY 8f-1)")
    ap.add_argument("--nchw", action="store_true", help="keep activations NCHW-contiguous (default: channels_last, "
                    "which spares cuDNN its NCHW<->NHWC transposes; the quantizer kernels are layout-agnostic)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
