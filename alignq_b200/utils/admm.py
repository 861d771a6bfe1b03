"""``ADMM(dim)`` loss module -- mirror of the reference's ``utils/admm.py:12-33`` on the CUDA kernel
``alignq_admm_loss`` (one 1024-thread CTA: three reductions, loss and all gradients in one launch)."""
from __future__ import annotations

import torch
from torch.nn.modules.module import Module
from torch.nn.parameter import Parameter

from .. import _lib as L


class _AdmmLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, D, alterD, gamma, mu, rho):
        Dc = L.dev_f32(D, "D")
        Z = L.dev_f32(alterD, "alterD")
        U = L.dev_f32(gamma, "gamma")
        B, dim = Dc.shape[0], Z.shape[0]
        if Dc.dim() != 2 or Dc.shape[1] != B or Z.shape != (dim, dim) or U.shape != (dim, dim) or dim < B:
            raise L.AlignQError(f"ADMM expects square D [B,B] with B <= dim; got D {tuple(D.shape)}, dim {dim}")
        loss = torch.empty((), dtype=torch.float32, device=Dc.device)
        with torch.cuda.device_of(Dc):
            L.check(L.load().alignq_admm_loss(Dc.data_ptr(), B, Z.data_ptr(), U.data_ptr(), dim, 1, mu, rho,
                                              0, 0, loss.data_ptr(), 0, 0, 0, L.stream_ptr()), "alignq_admm_loss")
        ctx.save_for_backward(Dc, Z, U)
        ctx.cfg = (mu, rho)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        Dc, Z, U = ctx.saved_tensors
        mu, rho = ctx.cfg
        B, dim = Dc.shape[0], Z.shape[0]
        gl = L.dev_f32(gloss.reshape(1), "grad of ADMM loss")
        gD = torch.empty_like(Dc) if ctx.needs_input_grad[0] else None
        gZ = torch.empty_like(Z) if ctx.needs_input_grad[1] else None
        gU = torch.empty_like(U) if ctx.needs_input_grad[2] else None
        with torch.cuda.device_of(Dc):
            L.check(L.load().alignq_admm_loss(Dc.data_ptr(), B, Z.data_ptr(), U.data_ptr(), dim, 1, mu, rho,
                                              gl.data_ptr(), 0, 0, L.ptr(gD), L.ptr(gZ), L.ptr(gU),
                                              L.stream_ptr()), "alignq_admm_loss (backward)")
        return gD, gZ, gU, None, None


class ADMM(Module):
    """ADMM loss: ``mu*mean|Z| + rho/2*sqrt(mean((D-Z)^2)) + mean(U*|D-Z|)`` with Z = alterD[:B,:B],
    U = gamma[:B,:B]; stores ``self.D`` for ``ADMM_OPT.step`` (admm.py:24-33)."""

    def __init__(self, dim):
        super().__init__()
        self.mu = 0.2
        self.rho = 0.3
        self.alterD = Parameter(torch.rand(dim, dim))
        self.gamma = Parameter(torch.rand(dim, dim))
        # d trans_loss / d(alterD, gamma) is produced by autograd in the reference but never read by ADMM_OPT
        # (closed-form Z/U updates, optimizer.py:97-124).  An AdmmBank that owns this module's update switches
        # it off PER MODULE (never through the process-global args) and serves D slots for any batch size.
        self.param_grads = True
        self._bank = None

    def d_slot(self, B: int):
        """Where the fused forward should write D for a batch of B rows (None: allocate)."""
        return None if self._bank is None else self._bank[0].slot(self._bank[1], int(B))

    def forward(self, D):
        self.D = D
        return _AdmmLossFn.apply(D, self.alterD, self.gamma, float(self.mu), float(self.rho))
