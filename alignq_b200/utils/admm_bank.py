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
        self.D = torch.zeros(n, self.B, self.B, dtype=torch.float32, device=dev)
        for i, m in enumerate(self.mods):
            self.Z[i].copy_(m.alterD.data)
            self.U[i].copy_(m.gamma.data)
            m.alterD.data = self.Z[i]
            m.gamma.data = self.U[i]
            m._D_slot = self.D[i]

    def ready(self) -> bool:
        """True when every module's last forward wrote its D into the bank (same batch size)."""
        return all(getattr(m, "D", None) is not None and m.D.data_ptr() == m._D_slot.data_ptr() for m in self.mods)

    @torch.no_grad()
    def update(self):
        """Z <- shrink(D + U/rho), U <- U + rho (D - Z) for every module in one launch."""
        with torch.cuda.device_of(self.Z):
            L.check(L.load().alignq_admm_zu_update(self.Z.data_ptr(), self.U.data_ptr(), self.D.data_ptr(), self.B,
                                                   self.dim, len(self.mods), self.mu, self.rho, L.stream_ptr()),
                    "alignq_admm_zu_update (bank)")
