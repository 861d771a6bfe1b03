"""Model-level golden vectors from the UNMODIFIED reference (run in the build container only).

    python oracle/make_model_golden.py

Per experiment directory (one subprocess each, SURVEY.md Appendix C shims) it records
  * the state_dict key order and shapes of the reference model      -> tests/golden/model_keys.json
  * logits / trans_loss / gradient fingerprints for seeded inputs and ``deterministic_fill`` weights,
    and the logits after two reference training iterations            -> tests/golden/model_<name>.npz
and asserts that ``oracle/models_oracle.py`` (ResNet-20, QA and QB) reproduces the reference
bit for bit on CPU, including two full training iterations with the reference's own SGD / ADMM_OPT.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
JOBS = {
    "resnet20_A": ("cdf_alignment/resnet-20-cifar-10", "model.resnet", "resnet20_quant", dict(bitW=8, abitW=8), 8, 32),
    "resnet20_B": ("cdf_alignment_admm/resnet-56-cifar-10", "model.resnet", "resnet20_quant", dict(bitW=8, abitW=8), 8, 32),
    "resnet56_B": ("cdf_alignment_admm/resnet-56-cifar-10", "model.resnet", "resnet56_quant", dict(bitW=8, abitW=8), 8, 32),
    "mobilenetv2_A": ("cdf_alignment/mobilenet-v2-svhn", "model.mobilenetV2", "mobile_v2", dict(wbit=4, abit=4), 4, 32),
    "densenet40_A": ("cdf_alignment/dense-cifar-10", "model.densenet", "densenet_40_quant", dict(bitW=8, abitW=8), 4, 32),
    "resnet50dann_C": ("cdf_alignment_admm/dann_office", "model.resnet", "resnet50_dann", dict(wbit=8, abit=8), 2, 96),
}


def fingerprint(model):
    """Per-parameter gradient fingerprint: (sum, abs-sum, first 4 values)."""
    import torch
    rows = []
    for _, p in model.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        f = g.detach().double().flatten()
        head = torch.zeros(4, dtype=torch.float64)
        head[: min(4, f.numel())] = f[:4]
        rows.append(torch.cat([f.sum().view(1), f.abs().sum().view(1), head]))
    return torch.stack(rows).numpy()


def worker(job: str) -> None:
    import torch
    exp, modname, ctor, kw, B, HW = JOBS[job]
    bits = kw.get("bitW", kw.get("wbit"))
    sys.argv = ["x", "--bitW", str(bits), "--abitW", str(bits), "--train_batch_size", str(B)]
    sys.path.insert(0, os.path.join(REF, exp))
    sys.path.insert(1, REPO)
    import importlib
    import model.quantization as q
    mod = importlib.import_module(modname)
    q.device = torch.device("cpu")
    if hasattr(mod, "device"):
        mod.device = torch.device("cpu")
    if hasattr(mod, "load_state_dict_from_url"):
        mod.load_state_dict_from_url = lambda *a, **k: {}
    q.args.act_range = 2
    q.args.method = "ours"
    from oracle import models_oracle as MO

    torch.manual_seed(0)
    ref = getattr(mod, ctor)(stage="second", **kw)
    ref.train()
    keys = {k: list(v.shape) for k, v in ref.state_dict().items()}
    pnames = [n for n, _ in ref.named_parameters()]
    sd = MO.deterministic_fill({k: v.clone() for k, v in ref.state_dict().items()}, seed=11)
    ref.load_state_dict(sd)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, 3, HW, HW, generator=g)
    ncls = 31 if "dann" in job else 10
    tgt = torch.randint(0, ncls, (B,), generator=g)
    out = {"x": x.numpy(), "target": tgt.numpy()}

    def run(model):
        r = model(x, 0.5) if "dann" in job else model(x)
        tl = None
        if isinstance(r, tuple):
            logits, tl = r[0], r[-1]
            if not torch.is_tensor(tl):          # 32-bit models return the python float 0
                tl = None
        else:
            logits = r
        loss = torch.nn.functional.cross_entropy(logits, tgt)
        total = loss if tl is None else loss + tl
        for p in model.parameters():
            p.grad = None
        total.backward()
        return logits.detach(), (None if tl is None else tl.detach()), fingerprint(model)

    logits, tl, fp = run(ref)
    out["logits"], out["grad_fp"] = logits.numpy(), fp
    if tl is not None:
        out["trans_loss"] = tl.numpy()

    # Sensitivity band of the REFERENCE itself: quantisation is discontinuous, so a 1-ulp change of the
    # weights flips codes and moves the outputs; a faithful re-implementation can only be asked to
    # stay inside (a small multiple of) this band at model level.
    with torch.no_grad():
        for p_ in ref.parameters():
            p_.mul_(1.0 + 1e-7)
    lp, tlp, fpp = run(ref)
    rel = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
    out["band_logits"] = np.float64(rel(lp, logits))
    out["band_grad_fp"] = np.float64(rel(torch.from_numpy(fpp[:, 1]), torch.from_numpy(fp[:, 1])))
    if tl is not None:
        out["band_trans_loss"] = np.float64(abs(float(tlp) - float(tl)) / abs(float(tl)))
    ref.load_state_dict(sd)

    # Wiring check without any quantizer: the same topology at 32 bit is a smooth function, so the
    # product graph (conv/bn/relu/shortcut order) can be compared tightly.
    kw32 = {k: 32 for k in kw}
    torch.manual_seed(0)
    ref32 = getattr(mod, ctor)(stage="second", **kw32)
    ref32.train()
    ref32.load_state_dict(sd)
    l32, _, fp32_ = run(ref32)
    out["logits_fp32"], out["grad_fp_fp32"] = l32.numpy(), fp32_
    print(f"  {job}: band logits {out['band_logits']:.2e}, grads {out['band_grad_fp']:.2e}")

    if job in ("resnet20_A", "resnet20_B", "resnet56_B"):
        variant = job[-1]
        units = [3, 3, 3] if "20" in job else [9, 9, 9]
        orc = MO.OracleResNet(units, 8, 8, variant, 2.0, dim=B)
        orc.load_state_dict(sd)
        orc.train()
        lo, tlo, fpo = run(orc)
        assert torch.equal(lo, logits), "oracle model logits != reference"
        assert tl is None or torch.equal(tlo, tl), "oracle model trans_loss != reference"
        assert np.array_equal(fpo, fp), "oracle model grads != reference"

        # two reference training iterations (main.py body) vs OracleTrainer
        import utils.optimizer as ropt
        ropt.args.bitW = 8
        ref.load_state_dict(sd)
        orc.load_state_dict(sd)
        named = [(n, p) for n, p in ref.named_parameters() if "alterD" not in n and "gamma" not in n]
        opt = ropt.SGD([p for _, p in named], lr=0.04, momentum=0.9, weight_decay=1e-4)
        admm_params = [(n, p) for n, p in ref.named_parameters() if "alterD" in n or "gamma" in n]
        opt_admm = ropt.ADMM_OPT([p for _, p in admm_params]) if variant != "A" else None
        trainer = MO.OracleTrainer(orc, lr=0.04, momentum=0.9, weight_decay=1e-4, lam=1.0, lam2=4.0, bitW=8)
        for it in range(2):
            opt.zero_grad()
            if opt_admm is not None:
                opt_admm.zero_grad()
            r = ref(x)
            if variant == "A":
                torch.nn.functional.cross_entropy(r, tgt).backward()
            else:
                o_t, t_l = r
                torch.nn.functional.cross_entropy(o_t, tgt).backward(retain_graph=True)
                t_l = t_l + 0.5
                t_l.backward()
            idx = [j for j, (n, _) in enumerate(named) if "conv" in n and "weight" in n][1:]
            if variant == "A":      # QA does not store the attributes: recompute them
                w_cdf, w_pdf = [], []
                for layer in ref.layers:
                    for conv in (layer.conv0, layer.conv1, layer.skip_conv):
                        if conv is not None:
                            c, p_ = q.cdf(torch.mean(conv.weight), torch.std(conv.weight), "w")(conv.weight)
                            w_cdf.append(c.detach())
                            w_pdf.append(p_.detach())
            else:
                w_cdf = [c.quantize_fn.weight_cdf for l in ref.layers for c in (l.conv0, l.conv1, l.skip_conv) if c is not None]
                w_pdf = [c.quantize_fn.weight_pdf for l in ref.layers for c in (l.conv0, l.conv1, l.skip_conv) if c is not None]
            opt.step(idx, w_cdf, w_pdf, 1.0, 4.0)
            if opt_admm is not None:
                a_idx = [j for j, (n, _) in enumerate(admm_params) if "alterD" in n]
                g_idx = [j for j, (n, _) in enumerate(admm_params) if "gamma" in n]
                mods = [ref.admm0]
                for l in ref.layers:
                    mods += [l.admm0, l.admm1] + ([l.admm_skip] if l.skip_conv is not None else [])
                opt_admm.step(a_idx, g_idx, [m.D for m in mods], [m.alterD for m in mods], [m.gamma for m in mods],
                              [m.mu for m in mods], [m.rho for m in mods])
            trainer.step(x, tgt)
            for (n, p), (_, po) in zip(ref.named_parameters(), orc.named_parameters()):
                assert torch.equal(p.detach(), po.detach()), f"iteration {it}: parameter {n} diverged"
        with torch.no_grad():
            r = ref(x)
            out["logits_after_2_steps"] = (r[0] if isinstance(r, tuple) else r).numpy()
        print(f"  {job}: oracle model + 2 training iterations bit-identical to the reference")

    if job == "densenet40_A":
        orc = MO.OracleDenseNet(8, 8, 2.0)
        orc.load_state_dict(sd)
        orc.train()
        lo, _, fpo = run(orc)
        assert torch.equal(lo, logits), "oracle DenseNet logits != reference"
        assert np.array_equal(fpo, fp), "oracle DenseNet grads != reference"
        # two reference training iterations (dense-cifar-10/main.py:285-325) vs OracleTrainer
        import utils.optimizer as ropt
        ropt.args.bitW = 8
        ref.load_state_dict(sd)
        orc.load_state_dict(sd)
        named = list(ref.named_parameters())
        opt = ropt.SGD([p for _, p in named], lr=0.04, momentum=0.9, weight_decay=1e-4)
        trainer = MO.OracleTrainer(orc, lr=0.04, momentum=0.9, weight_decay=1e-4, lam=1.0, lam2=4.0, bitW=8)
        for it in range(2):
            opt.zero_grad()
            torch.nn.functional.cross_entropy(ref(x), tgt).backward()
            idx = [j for j, (n, _) in enumerate(named) if "conv" in n and "weight" in n]        # main.py:297-300: all of them
            convs = [ref.conv1]
            for j, layer in enumerate([ref.dense1, ref.trans1, ref.dense2, ref.trans2, ref.dense3]):
                convs += [blk.conv1 for blk in layer] if j % 2 == 0 else [layer.conv1]
            w_cdf, w_pdf = [], []          # QA does not store the attributes: recompute them
            for conv in convs:
                c, p_ = q.cdf(torch.mean(conv.weight), torch.std(conv.weight), "w")(conv.weight)
                w_cdf.append(c.detach())
                w_pdf.append(p_.detach())
            opt.step(idx, w_cdf, w_pdf, 1.0, 4.0)
            trainer.step(x, tgt)
            for (n, p), (_, po) in zip(ref.named_parameters(), orc.named_parameters()):
                assert torch.equal(p.detach(), po.detach()), f"iteration {it}: parameter {n} diverged"
        with torch.no_grad():
            out["logits_after_2_steps"] = ref(x).numpy()
        print(f"  {job}: oracle DenseNet + 2 training iterations bit-identical to the reference")

    if job == "mobilenetv2_A":
        orc = MO.OracleMobileNetV2(4, 4, 2.0)
        orc.load_state_dict(sd)
        orc.train()
        lo, _, fpo = run(orc)
        assert torch.equal(lo, logits), "oracle MobileNet-v2 logits != reference"
        assert np.array_equal(fpo, fp), "oracle MobileNet-v2 grads != reference"
        # two reference training iterations (mobilenet-v2-svhn/main.py:160-204) vs OracleTrainer
        import utils.optimizer as ropt
        ropt.args.bitW = 4
        ref.load_state_dict(sd)
        orc.load_state_dict(sd)
        named = list(ref.named_parameters())
        opt = ropt.SGD([p for _, p in named], lr=0.04, momentum=0.9, weight_decay=1e-4)
        trainer = MO.OracleTrainer(orc, lr=0.04, momentum=0.9, weight_decay=1e-4, lam=1.0, lam2=4.0, bitW=4)
        for it in range(2):
            opt.zero_grad()
            torch.nn.functional.cross_entropy(ref(x), tgt).backward()
            idx = [j for j, (n, _) in enumerate(named)
                   if ("conv" in n and "weight" in n) or ("shortcut.0" in n and "weight" in n)]          # main.py:175-179
            convs = [ref.conv1]
            for layer in ref.layers:
                convs += [layer.conv1, layer.conv2, layer.conv3] + ([layer.shortcut[0]] if layer.shortcut is not None else [])
            convs.append(ref.conv2)
            w_cdf, w_pdf = [], []          # QA does not store the attributes: recompute them
            for conv in convs:
                c, p_ = q.cdf(torch.mean(conv.weight), torch.std(conv.weight), "w")(conv.weight)
                w_cdf.append(c.detach())
                w_pdf.append(p_.detach())
            opt.step(idx, w_cdf, w_pdf, 1.0, 4.0)
            trainer.step(x, tgt)
            for (n, p), (_, po) in zip(ref.named_parameters(), orc.named_parameters()):
                assert torch.equal(p.detach(), po.detach()), f"iteration {it}: parameter {n} diverged"
        with torch.no_grad():
            out["logits_after_2_steps"] = ref(x).numpy()
        print(f"  {job}: oracle MobileNet-v2 + 2 training iterations bit-identical to the reference")

    os.makedirs(os.path.join(REPO, "tests", "golden"), exist_ok=True)
    np.savez_compressed(os.path.join(REPO, "tests", "golden", f"model_{job}.npz"), **out)
    with open(os.path.join(REPO, "tests", "golden", f"model_keys_{job}.json"), "w") as f:
        json.dump({"state_dict": keys, "named_parameters": pnames}, f)
    print(f"{job}: {len(keys)} state_dict keys, logits {tuple(logits.shape)} written")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--job", choices=list(JOBS))
    a = ap.parse_args()
    if a.job:
        worker(a.job)
        return
    for j in JOBS:
        subprocess.run([sys.executable, os.path.abspath(__file__), "--job", j], check=True)


if __name__ == "__main__":
    main()
