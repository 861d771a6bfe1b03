"""Model-level batching of the ADMM dual variables: every ``ADMM`` module's ``alterD`` (Z) and ``gamma``
(U) become views of two stacked ``[L, dim, dim]`` buffers and every layer's ``D`` is written into its slot
of one ``[L, B, B]`` buffer, so ``ADMM_OPT.step`` (utils/optimizer.py:97-124; a Python loop with a host
sync per module in the reference) is ONE batched ``alignq_admm_zu_update`` launch (one 8-CTA cluster per
module).  Names, shapes and ``state_dict`` are unchanged.
"""
from __future__ import annotations

import torch

from .. import _lib as L
from .admm import ADMM


class AdmmBank:
    def __init__(self, model, batch):
        self.mods = [m for m in model.modules() if isinstance(m, ADMM)]
        if not self.mods:
            raise L.AlignQError("AdmmBank: the model has no ADMM modules")
        dims = {tuple(m.alterD.shape) for m in self.mods} | {tuple(m.gamma.shape) for m in self.mods}
        hp = {(float(m.mu), float(m.rho)) for m in self.mods}
        if len(dims) != 1 or len(hp) != 1:
            raise L.AlignQError("AdmmBank needs one dim and one (mu, rho) for all ADMM modules")
        self.dim = self.mods[0].alterD.shape[0]
        self.mu, self.rho = hp.pop()
        self.B = int(batch)
        if self.B > self.dim:
            raise L.AlignQError(f"batch {self.B} exceeds ADMM dim {self.dim}")
        dev = self.mods[0].alterD.device
        n = len(self.mods)
        self.Z = torch.empty(n, self.dim, self.dim, dtype=torch.float32, device=dev)
        self.U = torch.empty_like(self.Z)
        # D slots per batch size: [L, B, B]; the training batch is allocated now (graph capture must not
        # allocate), a ragged last batch (B < train_batch_size) gets its own set on first use
        self._D = {self.B: torch.zeros(n, self.B, self.B, dtype=torch.float32, device=dev)}
        for i, m in enumerate(self.mods):
            self.Z[i].copy_(m.alterD.data)
            self.U[i].copy_(m.gamma.data)
            m.alterD.data = self.Z[i]
            m.gamma.data = self.U[i]
            m._bank = (self, i)
            m.param_grads = False                 # per module: the bank applies the closed-form update
            m.alterD._alignq_closed_form = m.gamma._alignq_closed_form = True
            m.D = None

    @property
    def D(self):
        return self._D[self.B]

    def slot(self, i: int, B: int):
        if B > self.dim:
            raise L.AlignQError(f"batch {B} exceeds ADMM dim {self.dim}")
        buf = self._D.get(B)
        if buf is None:
            buf = self._D[B] = torch.zeros(len(self.mods), B, B, dtype=torch.float32, device=self.Z.device)
        return buf[i]

    def begin_iteration(self):
        """Forget last iteration's D handles: ready() must only see modules that ran in THIS forward."""
        for m in self.mods:
            m.D = None

    def ready(self):
        """Batch size B when every module's last forward wrote its D into the bank's [L, B, B] slots of ONE
        batch size (then update() covers all modules in one launch); 0 otherwise."""
        Bs = {int(m.D.shape[0]) if getattr(m, "D", None) is not None else -1 for m in self.mods}
        if len(Bs) != 1:
            return 0
        B = Bs.pop()
        buf = self._D.get(B)
        if B <= 0 or buf is None:
            return 0
        return B if all(m.D.data_ptr() == buf[i].data_ptr() for i, m in enumerate(self.mods)) else 0

    def release(self):
        for m in self.mods:
            m._bank = None
            m.param_grads = True
            for p in (m.alterD, m.gamma):
                if hasattr(p, "_alignq_closed_form"):
                    del p._alignq_closed_form

    @torch.no_grad()
    def update(self, B=None):
        """Z <- shrink(D + U/rho), U <- U + rho (D - Z) for every module in one launch (batch size B)."""
        B = self.B if B is None else int(B)
        with torch.cuda.device_of(self.Z):
            L.check(L.load().alignq_admm_zu_update(self.Z.data_ptr(), self.U.data_ptr(), self._D[B].data_ptr(), B,
                                                   self.dim, len(self.mods), self.mu, self.rho, L.stream_ptr()),
                    "alignq_admm_zu_update (bank)")
