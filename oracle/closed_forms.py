"""Closed-form fp64 gradients of the AlignQ hot path.

TEST INFRASTRUCTURE ONLY (see oracle/alignq_oracle.py header for the import rule).

These are a second opinion on what the reference's *autograd* computes; they
are checked against autograd of ``alignq_oracle`` in fp64 by
``tests/test_oracle_closed_forms.py`` and are the formulas the CUDA backward
kernels implement.  Each function cites the reference forward it differentiates.
"""
from __future__ import annotations

import math

import torch

SQRT_2PI = math.sqrt(2.0 * math.pi)


def phi(z: torch.Tensor) -> torch.Tensor:
    return torch.exp(-0.5 * z * z) / SQRT_2PI


def act_backward(x, gy, a_bit: int, variant: str, act_range: float, stage: str = "second"):
    """d/dx of activation_quantize_fn.forward (QA:91-103, QB:102-132) with the
    straight-through rounding (QA:29-32)."""
    if a_bit == 32 and stage != "align":
        return gy.clone()
    if variant == "A" and a_bit == 32:          # returns the bare CDF (QA:100-101)
        return gy * phi(x)
    return gy * (2.0 * act_range) * phi(x)


def weight_backward(w, g, w_bit: int):
    """d/dw of weight_quantize_fn.forward (QA:62-78 / QB:71-85): the gradient
    flows through torch.mean and torch.std (QA:70).  Same for QA and QB."""
    if w_bit == 32:
        return g.clone()
    N = w.numel()
    m = w.mean()
    s = w.std()
    z = (w - m) / s
    a = 2.0 * g * phi(z)
    sa = a.sum()
    saz = (a * z).sum()
    return (a - sa / N - z * saz / (N - 1)) / s


def admm_dloss_dD(D, Z, U, mu: float, rho: float):
    """d/dD of ADMM.forward (admm.py:24-33).  Z, U already sliced to D's shape."""
    R = D - Z
    B2 = D.numel()
    rms = torch.sqrt(torch.mean(R * R))
    return (rho / 2.0) * R / (B2 * rms) + U * torch.sign(R) / B2


def corr_backward(X, dG, eps: float):
    """d/dX of corr(X, X) (QB:134-137 / QC:158-161) for upstream dG [B,B]."""
    B, F = X.shape
    mu = X.mean(dim=0)
    sd = X.std(dim=0)
    c = X - mu
    Xs = c / (sd + eps)
    gS = (dG + dG.t()) @ Xs / F
    dsd = -(gS * c).sum(dim=0) / (sd + eps) ** 2
    through_sd = torch.where(sd > 0, dsd * c / ((B - 1) * sd), torch.zeros_like(c))   # torch masks std == 0
    return (gS - gS.mean(dim=0)) / (sd + eps) + through_sd


def act_admm_backward(X, gy, gloss, Z, U, mu, rho, act_range, eps):
    """Total d/dX of (y, trans_loss) = activation_quantize_fn(x) in QB/QC with
    upstream grads gy (for y) and the scalar gloss (for trans_loss).  X is [B,F]."""
    from . import alignq_oracle as O

    T = act_range * (2.0 * O.normal_cdf(X, torch.zeros(1, dtype=X.dtype),
                                         torch.ones(1, dtype=X.dtype)) - 1.0)
    D = O.corr(T, T, eps) - O.corr(X, X, eps)
    dD = admm_dloss_dD(D, Z[: D.shape[0], : D.shape[1]], U[: D.shape[0], : D.shape[1]], mu, rho) * gloss
    dphi = (2.0 * act_range) * phi(X)
    return corr_backward(X, -dD, eps) + corr_backward(T, dD, eps) * dphi + gy * dphi
