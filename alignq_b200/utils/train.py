"""One QAT training iteration, restating the body of the reference's ``train()`` loops on the
drop-in modules (cdf_alignment/resnet-20-cifar-10/main.py:269-313; cdf_alignment_admm/
resnet-56-cifar-10/main.py:286-379), B200-first:

  * gradients live in ONE flat fp32 buffer (every ``p.grad`` is a view), so ``zero_grad`` is one memset,
    the data-parallel exchange is ONE NCCL all-reduce, and every pointer the multi-tensor SGD kernel
    sees is static;
  * the whole iteration (forward, both backward passes, SGD.step, ADMM_OPT.step) can be captured
    in a CUDA graph and replayed: ResNet-20 at batch 128 is launch-bound, not bandwidth-bound
    (SURVEY.md 7.3), so replay removes the Python + launch overhead of ~600 small kernels.

Only what the hot path needs: no data loading, logging, checkpointing or LR schedule.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .admm import ADMM
from .optimizer import ADMM_OPT, SGD
from .options import args
from .sharding import allreduce_mean_


def quantized_convs(model):
    """Conv2d_Q / Linear_Q modules in registration order."""
    return [m for m in model.modules() if hasattr(m, "quantize_fn") and hasattr(m, "weight")]


def collect_sgd_args(model, params):
    """``idx, w_cdf, w_pdf`` as the reference gathers them before ``optimizer_t.step`` (main.py:295-309):
    every quantized conv weight except the stem (``idx = idx[1:]``), matched by parameter identity so
    it works for any of the model files; ``idx`` are positions in ``params``."""
    pos = {id(p): j for j, p in enumerate(params)}
    idx, w_cdf, w_pdf = [], [], []
    for conv in quantized_convs(model)[1:]:
        q = conv.quantize_fn
        if id(conv.weight) in pos and getattr(q, "weight_pdf", None) is not None and q.w_bit < 32:
            idx.append(pos[id(conv.weight)])
            w_cdf.append(q.weight_cdf)
            w_pdf.append(q.weight_pdf)
    return idx, w_cdf, w_pdf


def collect_admm_args(model, admm_params):
    """The seven lists of ``optimizer_admm.step`` (cdf_alignment_admm/.../main.py:328-374)."""
    mods = [m for m in model.modules() if isinstance(m, ADMM)]
    pos = {id(p): j for j, p in enumerate(admm_params)}
    mods = [m for m in mods if id(m.alterD) in pos and getattr(m, "D", None) is not None]
    mods.sort(key=lambda m: pos[id(m.alterD)])
    return ([pos[id(m.alterD)] for m in mods], [pos[id(m.gamma)] for m in mods], [m.D for m in mods],
            [m.alterD for m in mods], [m.gamma for m in mods], [m.mu for m in mods], [m.rho for m in mods])


class QATStep:
    """forward + backward + SGD.step (+ ADMM_OPT.step) on one batch; optionally graph-replayed and
    data-parallel (gradient all-reduce over the flat buffer; BN statistics via SyncBatchNorm are the
    caller's choice)."""

    def __init__(self, model, lr=0.04, momentum=0.9, weight_decay=1e-4, lam=None, lam2=None,
                 trans_loss_offset=0.5, process_group=None, world_size=1, bank_weights=True):
        self.model = model
        self.bank = None
        if bank_weights:                       # one multi-tensor weight-quantizer launch per step
            from .weight_bank import WeightBank
            try:
                self.bank = WeightBank(model)
            except Exception:
                self.bank = None
        named = list(model.named_parameters())
        self.params = [p for n, p in named if "alterD" not in n and "gamma" not in n]       # main.py:87
        self.admm_params = [p for n, p in named if "alterD" in n or "gamma" in n]
        self.opt = SGD(self.params, lr=lr, momentum=momentum, weight_decay=weight_decay)
        self.opt_admm = ADMM_OPT(self.admm_params) if self.admm_params else None
        self.lam = args.lam if lam is None else lam
        self.lam2 = args.lam2 if lam2 is None else lam2
        self.offset = trans_loss_offset
        self.pg, self.world = process_group, world_size
        dev = self.params[0].device
        allp = self.params + self.admm_params
        self.gflat = torch.zeros(sum(p.numel() for p in allp), dtype=torch.float32, device=dev)
        off = 0
        for p in allp:
            p.grad = self.gflat[off: off + p.numel()].view_as(p)
            off += p.numel()
        self.n_main = sum(p.numel() for p in self.params)
        self.graph = None
        self.static_x = self.static_t = None
        self.loss = torch.zeros((), device=dev)

    # -- one eager iteration -------------------------------------------------------------------
    def _iteration(self, x, t):
        self.gflat.zero_()                                         # optimizer.zero_grad(), one memset
        if self.bank is not None:
            self.bank.quantize_all()
        out = self.model(x)
        if isinstance(out, tuple):
            logits, trans_loss = out
            ce = F.cross_entropy(logits, t)
            if torch.is_tensor(trans_loss):
                ce.backward(retain_graph=True)                     # .../main.py:300-307
                (trans_loss + self.offset).backward()
            else:
                ce.backward()
        else:
            ce = F.cross_entropy(out, t)
            ce.backward()
        if self.world > 1:                                         # ONE collective per step over NVLink
            allreduce_mean_(self.gflat[: self.n_main], self.pg, self.world)
        idx, w_cdf, w_pdf = collect_sgd_args(self.model, self.params)
        self.opt.step(idx, w_cdf, w_pdf, self.lam, self.lam2)
        if self.opt_admm is not None:
            self.opt_admm.step(*collect_admm_args(self.model, self.admm_params))
        if self.bank is not None:
            self.bank.fresh = False                                # weights changed: slices are stale
        self.loss.copy_(ce.detach())
        return self.loss

    def step(self, x, t):
        if self.graph is None:
            return self._iteration(x, t)
        self.static_x.copy_(x, non_blocking=True)
        self.static_t.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.loss

    # -- CUDA-graph capture ----------------------------------------------------------------------
    def capture(self, x, t, warmup=3):
        """Warm up eagerly on a side stream (initialises momentum buffers, plans, workspaces), then
        capture one iteration.  ``x`` / ``t`` give the static shapes."""
        self.static_x, self.static_t = x.clone(), t.clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._iteration(self.static_x, self.static_t)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._iteration(self.static_x, self.static_t)
        self.graph = g
        return self
