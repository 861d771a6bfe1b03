"""Generate tests/golden/*.npz from the UNMODIFIED reference and pin the oracle.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_golden.py            # all three quantizer variants

For each variant (QA / QB / QC, SURVEY.md shorthand) a fresh subprocess imports
the reference experiment directory with the shims of SURVEY.md Appendix C
(argv before import, module-global ``device`` set to cpu), runs the reference
L1 modules on seeded inputs, asserts that ``oracle/alignq_oracle.py`` produces
BIT-IDENTICAL results on the same inputs (same torch, same CPU), and writes
inputs + reference outputs to ``tests/golden/l1_<variant>.npz``.  The GPU box
has no /root/reference; it checks the oracle against these files instead.
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DIRS = {
    "A": "cdf_alignment/resnet-20-cifar-10",
    "B": "cdf_alignment_admm/resnet-56-cifar-10",
    "C": "cdf_alignment_admm/dann_office",
}
BITS = (2, 4, 8)
ADMM_DIM = 8


def _same(a, b, what):
    import torch
    a = torch.as_tensor(a)
    b = torch.as_tensor(b)
    ok = a.shape == b.shape and bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all())
    if not ok:
        raise AssertionError(f"oracle != reference for {what}: max abs diff "
                             f"{(a.double() - b.double()).abs().max().item():.3e}")


def worker(variant: str) -> None:
    import torch
    sys.argv = ["x", "--bitW", "8", "--abitW", "8", "--train_batch_size", str(ADMM_DIM)]
    sys.path.insert(0, os.path.join(REF, DIRS[variant]))
    sys.path.insert(1, REPO)
    import model.quantization as q          # the reference, unmodified
    import utils.optimizer as ropt
    from utils.admm import ADMM as RefADMM
    from oracle import alignq_oracle as O
    from oracle import closed_forms as CF

    q.device = torch.device("cpu")
    q.args.act_range = 2
    q.args.method = "ours"
    ar = 2.0
    out = {}
    g = torch.Generator().manual_seed(1234 + ord(variant))

    def randn(*s, scale=1.0):
        return torch.randn(*s, generator=g) * scale

    # ---------------- activations (a1, a2, a4/a5, a8) ----------------------
    x = randn(ADMM_DIM, 6, 5, 5, scale=1.3)
    x.view(-1)[:6] = torch.tensor([0.0, -0.0, 7.5, -7.5, 1e-8, 38.0])   # edge inputs
    gy = randn(*x.shape)
    out["act_x"], out["act_gy"] = x.numpy(), gy.numpy()
    for k in BITS + (1, 32):
        stage = "align" if k == 32 else "second"
        if variant == "B":
            admm = RefADMM(ADMM_DIM)
            mod = q.activation_quantize_fn(k, stage, admm)
        else:
            mod = q.activation_quantize_fn(k, stage)
        xr = x.clone().requires_grad_(True)
        res = mod(xr)
        y = res[0] if isinstance(res, tuple) else res
        (y * gy).sum().backward()
        out[f"act_y_k{k}"], out[f"act_gx_k{k}"] = y.detach().numpy(), xr.grad.numpy()
        if k not in (1, 32):
            c_ref, _ = q.cdf(torch.zeros(1), torch.ones(1), "a")(x)
            codes = torch.round(c_ref * (2 ** k - 1))
            out[f"act_codes_k{k}"] = codes.numpy().astype(np.int32)
            _same(O.activation_codes(x, k, variant, ar), codes, f"act codes k={k}")
        # pin the oracle (forward bit-exact, autograd backward bit-exact)
        xo = x.clone().requires_grad_(True)
        yo = O.activation_quantize(xo, k, stage, variant, ar)
        (yo * gy).sum().backward()
        _same(yo.detach(), y.detach(), f"act y k={k}")
        _same(xo.grad, xr.grad, f"act gx k={k}")
        cf = CF.act_backward(x.double(), gy.double(), k, variant, ar, stage)
        assert torch.allclose(cf, xr.grad.double(), rtol=2e-5, atol=1e-9), f"closed form act bwd k={k}"

    # ---------------- weights (a3, a8) --------------------------------------
    w = randn(16, 8, 3, 3, scale=0.07) + 0.01
    gw_up = randn(*w.shape)
    out["w"], out["w_gup"] = w.numpy(), gw_up.numpy()
    for k in BITS + (32,):
        mod = q.weight_quantize_fn(k, "second")
        wr = w.clone().requires_grad_(True)
        wq = mod(wr)
        (wq * gw_up).sum().backward()
        out[f"w_q_k{k}"], out[f"w_g_k{k}"] = wq.detach().numpy(), wr.grad.numpy()
        wo = w.clone().requires_grad_(True)
        oq, oc, op = O.weight_quantize(wo, k, variant)
        (oq * gw_up).sum().backward()
        _same(oq.detach(), wq.detach(), f"wq k={k}")
        _same(wo.grad, wr.grad, f"gw k={k}")
        if k != 32:
            rc, rp = q.cdf(torch.mean(w), torch.std(w), "w")(w)
            out[f"w_cdf_k{k}"], out[f"w_pdf_k{k}"] = rc.numpy(), rp.numpy()
            _same(oc.detach(), rc, "weight_cdf")
            _same(op.detach(), rp, "weight_pdf")
            if variant != "A":
                _same(mod.weight_cdf.detach(), rc, "attr weight_cdf")
            cf = CF.weight_backward(w.double(), gw_up.double(), k)
            assert torch.allclose(cf, wr.grad.double(), rtol=1e-4, atol=1e-6), f"closed form w bwd k={k}"

    # ---------------- corr / ADMM / fused act+ADMM (a5, a6, a7, a8) ----------
    if variant != "A":
        eps = 0.0 if variant == "B" else 1e-5
        X = randn(ADMM_DIM, 96)
        Y = randn(ADMM_DIM, 96)
        G = q.corr(X, Y)
        out["corr_x"], out["corr_y"], out["corr_out"] = X.numpy(), Y.numpy(), G.numpy()
        _same(O.corr(X, Y, eps), G, "corr")

        admm = RefADMM(ADMM_DIM)
        with torch.no_grad():
            admm.alterD.copy_(torch.rand(ADMM_DIM, ADMM_DIM, generator=g))
            admm.gamma.copy_(torch.rand(ADMM_DIM, ADMM_DIM, generator=g))
        out["admm_Z"], out["admm_U"] = admm.alterD.detach().numpy(), admm.gamma.detach().numpy()
        for B in (ADMM_DIM, ADMM_DIM - 3):                  # full and ragged (B < dim) batches
            Dm = randn(B, B, scale=0.05).requires_grad_(True)
            loss = admm(Dm)
            loss.backward()
            out[f"admm_D_b{B}"], out[f"admm_loss_b{B}"] = Dm.detach().numpy(), loss.detach().numpy()
            out[f"admm_dD_b{B}"] = Dm.grad.numpy()
            _same(O.admm_loss(Dm.detach(), admm.alterD.detach(), admm.gamma.detach(), admm.mu, admm.rho),
                  loss.detach(), "admm loss")
            cf = CF.admm_dloss_dD(Dm.detach().double(), admm.alterD.detach()[:B, :B].double(),
                                  admm.gamma.detach()[:B, :B].double(), admm.mu, admm.rho)
            assert torch.allclose(cf, Dm.grad.double(), rtol=1e-5, atol=1e-9)

        Fn2 = q.activation_quantize_fn if variant == "B" else q.activation_quantize_fn2
        for B in (ADMM_DIM, ADMM_DIM - 3):
            for k in (4, 8):
                xa = x[:B].clone().requires_grad_(True)
                mod = Fn2(k, "second", admm)
                admm.zero_grad()
                y, tl = mod(xa)
                ((y * gy[:B]).sum() + 1.5 * tl).backward()
                tag = f"b{B}_k{k}"
                out[f"fused_y_{tag}"], out[f"fused_loss_{tag}"] = y.detach().numpy(), tl.detach().numpy()
                out[f"fused_D_{tag}"], out[f"fused_gx_{tag}"] = admm.D.detach().numpy(), xa.grad.numpy()
                xo = x[:B].clone().requires_grad_(True)
                yo, lo, Do = O.activation_quantize_admm(xo, k, admm.alterD.detach(), admm.gamma.detach(),
                                                        "second", variant, ar, admm.mu, admm.rho)
                ((yo * gy[:B]).sum() + 1.5 * lo).backward()
                _same(yo.detach(), y.detach(), "fused y")
                _same(lo.detach(), tl.detach(), "fused loss")
                _same(Do.detach(), admm.D.detach(), "fused D")
                _same(xo.grad, xa.grad, "fused gx")
                cf = CF.act_admm_backward(x[:B].double().view(B, -1), gy[:B].double().view(B, -1), 1.5,
                                          admm.alterD.detach().double(), admm.gamma.detach().double(),
                                          admm.mu, admm.rho, ar, eps).view_as(xa)
                assert torch.allclose(cf, xa.grad.double(), rtol=5e-3, atol=1e-6), "closed form fused bwd"

        # ------------- ADMM_OPT.step (a11) -----------------------------------
        for B in (ADMM_DIM, ADMM_DIM - 3):
            a2 = RefADMM(ADMM_DIM)
            with torch.no_grad():
                a2.alterD.copy_(admm.alterD)
                a2.gamma.copy_(admm.gamma)
            Dm = torch.from_numpy(out[f"admm_D_b{B}"])
            a2(Dm).backward()
            opt = ropt.ADMM_OPT([a2.alterD, a2.gamma])
            opt.step([0], [1], [Dm], [a2.alterD], [a2.gamma], [a2.mu], [a2.rho])
            out[f"zu_Z_b{B}"], out[f"zu_U_b{B}"] = a2.alterD.detach().numpy(), a2.gamma.detach().numpy()
            Zo, Uo = O.admm_zu_update(Dm, admm.alterD.detach(), admm.gamma.detach(), admm.mu, admm.rho)
            _same(Zo, a2.alterD.detach(), "Z update")
            _same(Uo, a2.gamma.detach(), "U update")
        # threshold branch: ||V|| <= mu/rho  ->  Z = 0
        a3 = RefADMM(ADMM_DIM)
        with torch.no_grad():
            a3.alterD.fill_(0.3)
            a3.gamma.fill_(1e-3)
        Dm = torch.full((ADMM_DIM, ADMM_DIM), 1e-3)
        a3(Dm).backward()
        ropt.ADMM_OPT([a3.alterD, a3.gamma]).step([0], [1], [Dm], [a3.alterD], [a3.gamma], [a3.mu], [a3.rho])
        out["zu_small_Z"], out["zu_small_U"] = a3.alterD.detach().numpy(), a3.gamma.detach().numpy()
        Zo, Uo = O.admm_zu_update(Dm, torch.full((ADMM_DIM, ADMM_DIM), 0.3),
                                  torch.full((ADMM_DIM, ADMM_DIM), 1e-3), 0.2, 0.3)
        _same(Zo, a3.alterD.detach(), "Z small")
        _same(Uo, a3.gamma.detach(), "U small")

    # ---------------- SGD.step (a10) ------------------------------------------
    ropt.args.bitW = 8
    shapes = [(4, 3, 3, 3), (6,), (8, 4, 3, 3), (8, 8, 1, 1), (5, 8)]
    idx = [2, 3]
    p0 = [randn(*s, scale=0.2) for s in shapes]
    grads = [[randn(*s, scale=0.05) for s in shapes] for _ in range(3)]
    wc = [torch.rand(*shapes[i], generator=g) * (2 if variant != "A" else 1) - (1 if variant != "A" else 0) for i in idx]
    wp = [torch.rand(*shapes[i], generator=g) * 3 for i in idx]
    for i, t in enumerate(p0):
        out[f"sgd_p0_{i}"] = t.numpy()
    for s_, gs in enumerate(grads):
        for i, t in enumerate(gs):
            out[f"sgd_g{s_}_{i}"] = t.numpy()
    for j in range(len(idx)):
        out[f"sgd_wcdf_{j}"], out[f"sgd_wpdf_{j}"] = wc[j].numpy(), wp[j].numpy()
    for cfg_name, kw in (("mom", dict(lr=0.04, momentum=0.9, weight_decay=1e-4)),
                         ("nest", dict(lr=0.02, momentum=0.8, weight_decay=5e-4, nesterov=True)),
                         ("plain", dict(lr=0.1))):
        ps = [torch.nn.Parameter(t.clone()) for t in p0]
        opt = ropt.SGD(ps, **kw)
        po = [t.clone() for t in p0]
        bufs = [None] * len(po)
        for s_, gs in enumerate(grads):
            for p, gg in zip(ps, gs):
                p.grad = gg.clone()
            opt.step(idx, wc, wp, 1.0, 4.0)
            og = O.sgd_step(po, [t.clone() for t in gs], bufs, idx, wc, wp, 1.0, 4.0, bitW=8,
                            **{"momentum": 0.0, "weight_decay": 0.0, **kw})
            for i, p in enumerate(ps):
                out[f"sgd_{cfg_name}_p{s_}_{i}"] = p.detach().numpy().copy()
                out[f"sgd_{cfg_name}_grad{s_}_{i}"] = p.grad.detach().numpy().copy()
                _same(po[i], p.detach(), f"sgd {cfg_name} p step{s_} #{i}")
                _same(og[i], p.grad.detach(), f"sgd {cfg_name} grad step{s_} #{i}")

    outdir = os.environ.get("ALIGNQ_GOLDEN_OUT") or os.path.join(REPO, "tests", "golden")   # tests regenerate into a temp dir
    os.makedirs(outdir, exist_ok=True)
    path = os.path.join(outdir, f"l1_{variant}.npz")
    np.savez_compressed(path, **out)
    print(f"variant {variant}: oracle == reference on {len(out)} arrays; wrote {path} "
          f"({os.path.getsize(path) / 1024:.0f} KiB)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", choices=list(DIRS))
    a = ap.parse_args()
    if a.variant:
        worker(a.variant)
        return
    for v in DIRS:          # one experiment dir per process (bare-name imports collide)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--variant", v], check=True)


if __name__ == "__main__":
    main()
