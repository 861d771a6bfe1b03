"""CPU: the oracle (oracle/alignq_oracle.py) against the golden vectors produced by the imported,
unmodified reference (oracle/make_golden.py).  Bit-exact on CPU: same torch ops, same order."""
import numpy as np
import pytest
import torch

from oracle import alignq_oracle as O
from oracle import closed_forms as CF

BITS = (2, 4, 8)


def t(a):
    return torch.from_numpy(np.asarray(a))


def same(a, b):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    return a.shape == b.shape and bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all())


@pytest.mark.parametrize("variant", ["A", "B", "C"])
def test_activation_forward_backward_bit_exact(golden, variant):
    g = golden(variant)
    x, gy = t(g["act_x"]), t(g["act_gy"])
    for k in BITS + (1, 32):
        stage = "align" if k == 32 else "second"
        xr = x.clone().requires_grad_(True)
        y = O.activation_quantize(xr, k, stage, variant, 2.0)
        (y * gy).sum().backward()
        assert same(y.detach(), t(g[f"act_y_k{k}"])), (variant, k)
        assert same(xr.grad, t(g[f"act_gx_k{k}"])), (variant, k)
        if k in BITS:
            assert same(O.activation_codes(x, k, variant, 2.0).to(torch.int32), t(g[f"act_codes_k{k}"]))


@pytest.mark.parametrize("variant", ["A", "B", "C"])
def test_weight_forward_backward_bit_exact(golden, variant):
    g = golden(variant)
    w, gup = t(g["w"]), t(g["w_gup"])
    for k in BITS + (32,):
        wr = w.clone().requires_grad_(True)
        wq, wc, wp = O.weight_quantize(wr, k, variant)
        (wq * gup).sum().backward()
        assert same(wq.detach(), t(g[f"w_q_k{k}"]))
        assert same(wr.grad, t(g[f"w_g_k{k}"]))
        if k != 32:
            assert same(wc.detach(), t(g[f"w_cdf_k{k}"]))
            assert same(wp.detach(), t(g[f"w_pdf_k{k}"]))


@pytest.mark.parametrize("variant", ["B", "C"])
def test_corr_admm_fused_bit_exact(golden, variant):
    g = golden(variant)
    eps = 0.0 if variant == "B" else 1e-5
    assert same(O.corr(t(g["corr_x"]), t(g["corr_y"]), eps), t(g["corr_out"]))
    Z, U = t(g["admm_Z"]), t(g["admm_U"])
    dim = Z.shape[0]
    for B in (dim, dim - 3):
        assert same(O.admm_loss(t(g[f"admm_D_b{B}"]), Z, U), t(g[f"admm_loss_b{B}"]))
        Zn, Un = O.admm_zu_update(t(g[f"admm_D_b{B}"]), Z, U)
        assert same(Zn, t(g[f"zu_Z_b{B}"])) and same(Un, t(g[f"zu_U_b{B}"]))
        for k in (4, 8):
            tag = f"b{B}_k{k}"
            x = t(g["act_x"])[:B].clone().requires_grad_(True)
            y, loss, D = O.activation_quantize_admm(x, k, Z, U, "second", variant, 2.0)
            ((y * t(g["act_gy"])[:B]).sum() + 1.5 * loss).backward()
            assert same(y.detach(), t(g[f"fused_y_{tag}"]))
            assert same(loss.detach(), t(g[f"fused_loss_{tag}"]))
            assert same(D.detach(), t(g[f"fused_D_{tag}"]))
            assert same(x.grad, t(g[f"fused_gx_{tag}"]))
    Zs, Us = O.admm_zu_update(torch.full((dim, dim), 1e-3), torch.full((dim, dim), 0.3), torch.full((dim, dim), 1e-3))
    assert same(Zs, t(g["zu_small_Z"])) and same(Us, t(g["zu_small_U"]))
    assert float(Zs.abs().max()) == 0.0            # the ||V|| <= mu/rho branch


@pytest.mark.parametrize("variant", ["A", "B"])
def test_sgd_step_bit_exact(golden, variant):
    g = golden(variant)
    idx = [2, 3]
    n = 5
    wc = [t(g[f"sgd_wcdf_{j}"]) for j in range(2)]
    wp = [t(g[f"sgd_wpdf_{j}"]) for j in range(2)]
    cfgs = {"mom": dict(lr=0.04, momentum=0.9, weight_decay=1e-4),
            "nest": dict(lr=0.02, momentum=0.8, weight_decay=5e-4, nesterov=True),
            "plain": dict(lr=0.1)}
    for name, kw in cfgs.items():
        ps = [t(g[f"sgd_p0_{i}"]).clone() for i in range(n)]
        bufs = [None] * n
        for s in range(3):
            grads = [t(g[f"sgd_g{s}_{i}"]).clone() for i in range(n)]
            og = O.sgd_step(ps, grads, bufs, idx, wc, wp, 1.0, 4.0, bitW=8, **kw)
            for i in range(n):
                assert same(ps[i], t(g[f"sgd_{name}_p{s}_{i}"])), (name, s, i)
                assert same(og[i], t(g[f"sgd_{name}_grad{s}_{i}"])), (name, s, i)


def test_closed_forms_match_oracle_autograd_fp64():
    """The formulas the CUDA backward kernels implement vs autograd of the oracle."""
    gen = torch.Generator().manual_seed(7)
    x = torch.randn(6, 40, generator=gen, dtype=torch.float64)
    gy = torch.randn(6, 40, generator=gen, dtype=torch.float64)
    for variant in ("A", "B"):
        xr = x.clone().requires_grad_(True)
        (O.activation_quantize(xr, 4, "second", variant, 2.0) * gy).sum().backward()
        assert torch.allclose(CF.act_backward(x, gy, 4, variant, 2.0), xr.grad, rtol=1e-12, atol=1e-14)
    w = torch.randn(300, generator=gen, dtype=torch.float64) * 0.1
    g = torch.randn(300, generator=gen, dtype=torch.float64)
    for variant in ("A", "B"):
        wr = w.clone().requires_grad_(True)
        (O.weight_quantize(wr, 4, variant)[0] * g).sum().backward()
        assert torch.allclose(CF.weight_backward(w, g, 4), wr.grad, rtol=1e-10, atol=1e-12)
    Z = torch.rand(6, 6, generator=gen, dtype=torch.float64)
    U = torch.rand(6, 6, generator=gen, dtype=torch.float64)
    for variant, eps in (("B", 0.0), ("C", 1e-5)):
        xr = x.clone().requires_grad_(True)
        y, loss, D = O.activation_quantize_admm(xr, 4, Z, U, "second", variant, 2.0)
        ((y * gy).sum() + 0.7 * loss).backward()
        cf = CF.act_admm_backward(x, gy, 0.7, Z, U, 0.2, 0.3, 2.0, eps)
        assert torch.allclose(cf, xr.grad, rtol=1e-9, atol=1e-12)
