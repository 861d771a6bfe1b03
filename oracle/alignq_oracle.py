"""CPU/torch-eager restatement of AlignQ's per-layer quantization hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``alignq_b200/`` may import this file.
Allowed importers: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs, where it is the *checker* (or the
timed CPU baseline), never the product.

Parity status: PINNED.  The reference (tinganchen/AlignQ) ships no tests and no
golden vectors, so the pin is the reference itself, imported
and executed in the build container by ``oracle/make_golden.py``; that script
asserts bit-equality between every function below and the reference module on
CPU, and writes the inputs/outputs to ``tests/golden/*.npz``.  The GPU box has no
``/root/reference``; there the checks are (1) this file against the committed
golden vectors and (2) the CUDA kernels against this file on the same inputs.

The arithmetic of this path lives in a third-party dependency of the
reference, PyTorch (pinned ``pytorch==1.7.1`` in README.md:9; 2.11 here):
``torch.distributions.Normal.cdf/log_prob``, ``torch.erf``, ``torch.round``,
``torch.mean/std``, ``torch.matmul``, ``torch.norm``.  The functions below spell out
those published formulas with the same torch primitives in the same order, so
the file is device-agnostic: run on CPU tensors it is the CPU oracle, run on
CUDA tensors it is "the reference's own PyTorch path on GPU-eager" that
BASELINE.json pins bit-exact codes against.

Reference files (relative to /root/reference):
  QA = cdf_alignment/*/model/quantization.py
  QB = cdf_alignment_admm/resnet-{20,56}-cifar-10/model/quantization.py
  QC = cdf_alignment_admm/{dann,dsan}_office/model/quantization.py
  admm.py / optimizer.py = */utils/{admm,optimizer}.py
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch

VARIANTS = ("A", "B", "C")


# --------------------------------------------------------------------------- #
# a1  uniform_quantize(k)                                   QA:15-34, QB:19-38 #
# --------------------------------------------------------------------------- #
class _RoundSTE(torch.autograd.Function):
    """round(x*n)/n forward, identity backward (QA:25-32)."""

    @staticmethod
    def forward(ctx, x, k):
        if k == 32:
            return x
        if k == 1:
            return torch.sign(x)
        n = 2 ** k - 1
        return torch.round(x * n) / n

    @staticmethod
    def backward(ctx, g):
        return g.clone(), None


def uniform_quantize(x: torch.Tensor, k: int) -> torch.Tensor:
    return _RoundSTE.apply(x, k)


def codes_of(x: torch.Tensor, k: int) -> torch.Tensor:
    """Integer codes round(x*n) the reference holds as fp32 (QA:26)."""
    n = 2 ** k - 1
    return torch.round(x * n)


# --------------------------------------------------------------------------- #
# a2  cdf(m, s, src)                                        QA:37-50, QB:41-59 #
# --------------------------------------------------------------------------- #
def normal_cdf(x: torch.Tensor, loc: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """torch.distributions.Normal(loc, scale).cdf(x) spelled out (called at QA:46-47)."""
    return 0.5 * (1 + torch.erf((x - loc) * scale.reciprocal() / math.sqrt(2)))


def normal_log_prob(x: torch.Tensor, loc: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """torch.distributions.Normal(loc, scale).log_prob(x) spelled out (called at QA:49)."""
    var = scale ** 2
    return -((x - loc) ** 2) / (2 * var) - scale.log() - math.log(math.sqrt(2 * math.pi))


def cdf_map(x, loc, scale, src: str, variant: str, act_range: float):
    """Returns (mapped, pdf).  QA: mapped = Phi.  QB/QC: 2*Phi-1 (times act_range for 'a')."""
    c = normal_cdf(x, loc, scale)
    if variant != "A":
        c = c * 2 - 1                                   # QB:53
        if src == "a":
            c = c * act_range                           # QB:55-56
    pdf = torch.exp(normal_log_prob(x, loc, scale)) * 2  # QA:49 / QB:58
    return c, pdf


# --------------------------------------------------------------------------- #
# a3  weight_quantize_fn.forward                            QA:62-78, QB:71-85 #
# --------------------------------------------------------------------------- #
def weight_quantize(w: torch.Tensor, w_bit: int, variant: str = "A"):
    """Returns (wq, weight_cdf, weight_pdf); the latter two are None for w_bit==32."""
    if w_bit == 32:
        return w, None, None
    wc, wp = cdf_map(w, torch.mean(w), torch.std(w), "w", variant, 0.0)
    if variant == "A":
        wq = uniform_quantize(wc, w_bit) * 2 - 1        # QA:72
    else:
        wq = uniform_quantize(wc, w_bit)                # QB:80
    return wq, wc, wp


# --------------------------------------------------------------------------- #
# a4/a5  activation_quantize_fn[2].forward     QA:91-103, QB:102-132, QC:96-156 #
# --------------------------------------------------------------------------- #
def activation_map(x: torch.Tensor, variant: str, act_range: float) -> torch.Tensor:
    zero = torch.zeros(1).to(x.device)                  # QA:97
    one = torch.ones(1).to(x.device)
    c, _pdf = cdf_map(x, zero, one, "a", variant, act_range)
    return c


def activation_quantize(x, a_bit: int, stage: str = "second", variant: str = "A",
                        act_range: float = 2.0) -> torch.Tensor:
    if a_bit == 32 and stage != "align":
        return x
    c = activation_map(x, variant, act_range)
    if variant == "A":
        q = (uniform_quantize(c, a_bit) * 2 - 1) * act_range   # QA:98
    else:
        q = uniform_quantize(c, a_bit)                         # QB:110
    return c if a_bit == 32 else q


def activation_codes(x, a_bit: int, variant: str = "A", act_range: float = 2.0):
    return codes_of(activation_map(x, variant, act_range), a_bit)


# --------------------------------------------------------------------------- #
# a6  corr(x, y)                                     QB:134-137, QC:158-161    #
# --------------------------------------------------------------------------- #
def corr(x: torch.Tensor, y: torch.Tensor, eps: float = 0.0) -> torch.Tensor:
    if eps == 0.0:
        xs = (x - torch.mean(x, dim=0)) / torch.std(x, dim=0)
        ys = (y - torch.mean(y, dim=0)) / torch.std(y, dim=0)
    else:
        xs = (x - torch.mean(x, dim=0)) / (torch.std(x, dim=0) + eps)
        ys = (y - torch.mean(y, dim=0)) / (torch.std(y, dim=0) + eps)
    return torch.matmul(xs, torch.transpose(ys, 0, 1)) / xs.shape[1]


# --------------------------------------------------------------------------- #
# a7  ADMM.forward(D)                                          admm.py:24-33   #
# --------------------------------------------------------------------------- #
def admm_loss(D, alterD, gamma, mu: float = 0.2, rho: float = 0.3):
    Z = alterD[: D.shape[0], : D.shape[1]]
    U = gamma[: D.shape[0], : D.shape[1]]
    reg = mu * torch.mean(torch.abs(Z))
    constraint = rho / 2 * torch.mean((D - Z) ** 2) ** 0.5
    relax = torch.mean(U * torch.abs(D - Z))
    return reg + constraint + relax


def activation_quantize_admm(x, a_bit: int, alterD, gamma, stage="second", variant="B",
                             act_range=2.0, mu=0.2, rho=0.3, method="ours"):
    """QB activation_quantize_fn / QC activation_quantize_fn2: returns (y, trans_loss, D)."""
    if a_bit == 32 and stage != "align":
        return x, 0, None
    eps = 0.0 if variant == "B" else 1e-5
    c = activation_map(x, variant, act_range)
    q = uniform_quantize(c, a_bit)
    D, loss = None, 0
    if method == "ours" and a_bit < 32:
        xf = x.view(x.shape[0], -1)
        tf = c.view(x.shape[0], -1)
        D = corr(tf, tf, eps) - corr(xf, xf, eps)       # QB:118-122
        loss = admm_loss(D, alterD, gamma, mu, rho)     # QB:123
    return (c if a_bit == 32 else q), loss, D


# --------------------------------------------------------------------------- #
# a10  SGD.step(idx, w_cdf, w_pdf, lam, lam2)         optimizer.py:196-262     #
# --------------------------------------------------------------------------- #
def _sigmoid(x):
    return 1 / (1 + torch.exp(-x))                      # optimizer.py:6-7


def surrogate(weight_cdf, lam, lam2, bitW):
    """sigmoid_d(transform(w_cdf, lam2), lam)           optimizer.py:9-13."""
    t = (((weight_cdf + 0.5) * (2 ** bitW - 1)) % 1) * lam2 * 2
    return _sigmoid(t) * (1 - _sigmoid(t)) * lam


def sgd_step(params: Sequence[torch.Tensor], grads: Sequence[Optional[torch.Tensor]],
             bufs: List[Optional[torch.Tensor]], idx: Sequence[int], w_cdf, w_pdf,
             lam: float, lam2: float, *, lr: float, momentum: float = 0.0,
             dampening: float = 0.0, weight_decay: float = 0.0, nesterov: bool = False,
             bitW: int = 8):
    """Functional restatement on plain tensors.  Mutates params/bufs in place and
    returns the list of tensors the reference leaves in ``p.grad``."""
    out_grads = []
    idx = list(idx)
    for i, (p, g) in enumerate(zip(params, grads)):
        if g is None:
            out_grads.append(None)
            continue
        d_p = g
        if weight_decay != 0:
            d_p.add_(p, alpha=weight_decay)             # in place on p.grad (:217)
        if momentum != 0:
            if bufs[i] is None:
                bufs[i] = torch.zeros_like(p)
                bufs[i].mul_(momentum).add_(d_p)
            else:
                bufs[i].mul_(momentum).add_(d_p, alpha=1 - dampening)
            d_p = d_p.add(bufs[i], alpha=momentum) if nesterov else bufs[i]
        if bitW < 32 and i in idx:
            j = idx.index(i)
            new_g = d_p * surrogate(w_cdf[j], lam, lam2, bitW) * w_pdf[j]
            p.add_(d_p, alpha=-lr)                      # update uses d_p, not new_g (:249-251)
            out_grads.append(new_g)
        else:
            p.add_(d_p, alpha=-lr)
            out_grads.append(d_p.clone())
    return out_grads


# --------------------------------------------------------------------------- #
# a11  ADMM_OPT.step(...)                               optimizer.py:60-135    #
# --------------------------------------------------------------------------- #
def admm_zu_update(D, alterD, gamma, mu: float = 0.2, rho: float = 0.3):
    """One (alterD, gamma) pair of ADMM_OPT.step; returns the new (Z, U)."""
    D_ = torch.zeros_like(gamma)
    D_[: D.shape[0], : D.shape[1]] = D
    V = D_ + 1 / rho * gamma
    nv = torch.norm(V, 2)
    if nv > (mu / rho):
        Z = (1 - mu / rho / nv) * V
    else:
        Z = torch.zeros_like(alterD)
    U = gamma + rho * (D_ - Z)
    return Z, U
