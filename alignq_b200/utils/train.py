"""One QAT training iteration, restating the body of the reference's ``train()`` loops on the
drop-in modules (cdf_alignment/resnet-20-cifar-10/main.py:269-313; cdf_alignment_admm/
resnet-56-cifar-10/main.py:286-379), B200-first:

  * ``zero_grad`` is ``p.grad = None`` (autograd hands over its buffers, no accumulate kernels); the
    data-parallel exchange all-reduces the weight bank's flat gradient buffer in place and gathers only the
    few small remaining gradients into a second flat buffer (two NCCL all-reduces per step, no copy of the
    weights' gradients); the 1/world mean is folded into the multi-tensor SGD kernel, which reads everything
    through a device pointer table;
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
from .. import _lib as L
from .sharding import allreduce_sum_


def WeightBank_applies(model) -> bool:
    """The bank needs at least one quantized conv / linear and ONE bit-width and variant for all of them."""
    qs = [m.quantize_fn for m in model.modules() if hasattr(m, "quantize_fn") and hasattr(m, "weight")
          and getattr(m.quantize_fn, "w_bit", 32) < 32]
    return bool(qs) and len({(q.w_bit, q.variant) for q in qs}) == 1


def single_backward_ok(step) -> bool:
    """The side-stream all-reduce of the upstream weight gradients assumes ONE backward pass per step (with two, e.g.
    CE.backward(retain_graph=True) then trans_loss.backward(), a bucket would be reduced after the first pass only)."""
    has_admm = bool(step.admm_params)
    return (not has_admm) or step.single_backward


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
                 trans_loss_offset=0.5, process_group=None, world_size=1, bank_weights=True,
                 channels_last=False, fast_admm=True, single_backward=False, forward_loss=None, keep_logits=False,
                 async_wgrad=True):
        self.model = model
        if channels_last:                      # NHWC weights: cuDNN needs no layout conversion kernels
            for p in model.parameters():
                if p.dim() == 4:
                    p.data = p.data.contiguous(memory_format=torch.channels_last)
        self.bank = None
        if bank_weights and WeightBank_applies(model):   # one multi-tensor weight-quantizer launch per step
            from .weight_bank import WeightBank
            self.bank = WeightBank(model)      # a failure here is an error, not a silent 2x slower step
            self.bank.batched_backward = True
        named = list(model.named_parameters())
        self.params = [p for n, p in named if "alterD" not in n and "gamma" not in n]       # main.py:87
        self.admm_params = [p for n, p in named if "alterD" in n or "gamma" in n]
        self.opt = SGD(self.params, lr=lr, momentum=momentum, weight_decay=weight_decay)
        self.opt_admm = ADMM_OPT(self.admm_params) if self.admm_params else None
        self.admm_bank = None
        if self.admm_params and fast_admm:     # batched Z/U update; d loss / d(Z, U) is never read by ADMM_OPT
            from .admm_bank import AdmmBank
            # per-module switch (ADMM.param_grads), never the process-global args; serves any batch size <= dim
            self.admm_bank = AdmmBank(model, args.train_batch_size)
        self.lam = args.lam if lam is None else lam
        self.lam2 = args.lam2 if lam2 is None else lam2
        self.offset = trans_loss_offset
        # reference: CE.backward(retain_graph=True) then trans_loss.backward() (resnet-56 main.py:300-307);
        # cdf_alignment_admm/resnet-20-cifar-10/main.py:297-300 does ONE (CE + trans_loss).backward(): same sums
        self.single_backward = single_backward
        # optional custom forward: (model, x, t) -> (task_loss, trans_loss | None), e.g. the two DANN passes of
        # cdf_alignment_admm/dann_office/main.py:372-385
        self.forward_loss = forward_loss
        self.keep_logits, self.logits = keep_logits, None       # the training driver reads the step's logits (accuracy)
        # conv weight gradients on a side stream (model/conv_tc.py): only with the bank's batched backward, the one
        # consumer of those gradients, which runs after the join in _backward()
        self.async_wgrad = bool(async_wgrad and self.bank is not None)
        self.pg, self.world = process_group, world_size
        # data parallel + side-stream weight gradients: the bank's upstream gradients are all-reduced on the side stream,
        # bucket by bucket, while the backward chain is still running (WeightBank.enable_dp_overlap)
        # NOT together with the in-kernel peer exchange of the BatchNorm sums: parallel branches of a CUDA graph (and
        # streams sharing a hardware queue) may be serialised by the driver, and then rank A's NCCL kernel, queued in front
        # of its BatchNorm kernel k, waits for rank B's NCCL kernel, which is queued behind B's BatchNorm kernel k that
        # spins for A's: a cross-GPU deadlock (seen on 2 GPUs: the bounded spin trapped).  Kernels that wait for another
        # GPU must all sit in ONE stream order.
        self.dp_overlap = bool(self.world > 1 and self.async_wgrad and single_backward_ok(self) and args.sync_bn != "peer")
        if self.dp_overlap:
            self.bank.enable_dp_overlap(self.pg, self.world)
        self.all_params = self.params + self.admm_params
        dev = self.params[0].device
        self.graph = None
        self.static_x = self.static_t = None
        self.loss = torch.zeros((), device=dev)

    # -- one eager iteration -------------------------------------------------------------------
    def _iteration(self, x, t):
        args.async_wgrad = self.async_wgrad                        # read by Conv2d_Q.forward of THIS step's model
        try:
            return self._iteration_body(x, t)
        finally:
            args.async_wgrad = False

    def _iteration_body(self, x, t):
        for p in self.all_params:                                  # optimizer.zero_grad(): autograd then hands
            p.grad = None                                          # over its gradient buffers without an add
        if self.bank is not None:
            self.bank.quantize_all()
        if self.admm_bank is not None:
            self.admm_bank.begin_iteration()
        if self.forward_loss is not None:
            ce, trans_loss = self.forward_loss(self.model, x, t)
        else:
            if args.fused_head and hasattr(self.model, "forward_ce"):
                ce, logits, trans_loss = self.model.forward_ce(x, t)       # pool -> linear -> loss: two launches
            else:
                out = self.model(x)
                logits, trans_loss = out if isinstance(out, tuple) else (out, None)
                ce = F.cross_entropy(logits, t)
            if self.keep_logits:
                if self.logits is None or self.logits.shape != logits.shape:
                    self.logits = torch.empty_like(logits)
                self.logits.copy_(logits.detach())
        if torch.is_tensor(trans_loss) and self.world > 1 and args.dp_gram == "feature":
            # global-batch trans_loss is the SAME scalar on every rank and each rank holds the gradient of its own
            # rows only: the mean over ranks (grad_scale = 1/world below) must see it world times
            trans_loss = trans_loss * float(self.world)
        if torch.is_tensor(trans_loss) and self.single_backward:
            self._backward(ce + trans_loss + self.offset)
        elif torch.is_tensor(trans_loss):
            self._backward(ce, retain_graph=True)                  # .../main.py:300-307
            self._backward(trans_loss + self.offset)
        else:
            self._backward(ce)
        scale = 1.0
        if self.world > 1:                                         # gradient exchange over NVLink
            owners = [p for p in self.params if p.grad is not None]
            if self.bank is not None and self.bank.batched_backward:
                # the weight bank's gradients already ARE one flat buffer (gw_flat; p.grad are views of it): it is
                # all-reduced in place, and only the few small remaining gradients (BatchNorm, first conv, classifier)
                # are gathered -- no torch.cat over the model's weights, no re-pointing of their p.grad
                banked = {id(p) for p, g in zip(self.bank.params, self.bank.gw) if p.grad is g}
                if not self.dp_overlap:                            # (overlap mode: already summed over the ranks on the side
                    allreduce_sum_(self.bank.gw_flat, self.pg)     #  stream, before the quantizer backward, which is linear)
                owners = [p for p in owners if id(p) not in banked]
            for p in owners:                                       # gather (physical order) -> all-reduce -> views
                if p.grad.stride() != p.stride():
                    p.grad = torch.empty_like(p).copy_(p.grad)
            if owners:
                flat = torch.cat([L.phys(p.grad) for p in owners])
                allreduce_sum_(flat, self.pg)
                off = 0
                for p in owners:
                    p.grad = torch.as_strided(flat, p.shape, p.stride(), off)
                    off += p.numel()
            scale = 1.0 / self.world                               # the mean is folded into the SGD kernel
        idx, w_cdf, w_pdf = collect_sgd_args(self.model, self.params)
        self.opt.step(idx, w_cdf, w_pdf, self.lam, self.lam2, grad_scale=scale)
        nb = self.admm_bank.ready() if self.admm_bank is not None else 0
        if nb:
            self.admm_bank.update(nb)                              # all modules, one launch (any batch size <= dim)
        elif self.opt_admm is not None:
            # some module did not run, or ran with a different batch: per-module updates of exactly the modules
            # that produced a D in THIS forward (ADMM_OPT treats bank-owned Z/U as present without .grad)
            self.opt_admm.step(*collect_admm_args(self.model, self.admm_params))
        if self.bank is not None:
            self.bank.fresh = False                                # weights changed: slices are stale
        self.loss.copy_(ce.detach())
        return self.loss

    def _backward(self, loss, retain_graph=False):
        loss.backward(retain_graph=retain_graph)
        if self.async_wgrad:
            from ..model.conv_tc import join_wgrads
            join_wgrads(self.bank if self.dp_overlap else None)    # side-stream weight gradients -> visible to this stream
        if self.bank is not None:
            self.bank.flush_backward()                             # all weight-quantizer backwards, one launch pair

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
        self.opt.defer_uploads_in_capture = True
        with torch.cuda.graph(g):
            self._iteration(self.static_x, self.static_t)
        self.opt.flush_deferred_uploads()                          # pointer tables of the captured update: once, not per replay
        self.graph = g
        return self


class HostFeeder:
    """Double-buffered host -> device input staging for ``QATStep``: the H2D copy of batch i+1 runs on its own
    stream while step i computes, so a training loop that reads its batches from (pinned) host memory does not
    serialise a copy in front of every step.  Usage::

        feeder = HostFeeder(step, x_example, t_example)
        feeder.prefetch(x0, t0)
        for (x_next, t_next) in batches:           # host tensors, ideally pinned
            loss = feeder.step(x_next, t_next)      # runs the staged batch, stages the next one meanwhile

    Every batch still crosses PCIe/NVLink-C2C exactly once; only the ordering changes."""

    def __init__(self, step: QATStep, x_like, t_like):
        dev = step.params[0].device
        fmt = torch.channels_last if (x_like.dim() == 4 and x_like.is_contiguous(memory_format=torch.channels_last)
                                      and not x_like.is_contiguous()) else torch.contiguous_format
        self.qat = step
        self.xs = [torch.empty(x_like.shape, dtype=x_like.dtype, device=dev).contiguous(memory_format=fmt) for _ in range(2)]
        self.ts = [torch.empty(t_like.shape, dtype=t_like.dtype, device=dev) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]
        self.slot = 0                      # slot holding the batch the next step() consumes
        self.staged = False

    def prefetch(self, host_x, host_t, slot=None):
        slot = self.slot if slot is None else slot
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[slot])       # the step that last read this slot is done with it
            self.xs[slot].copy_(host_x, non_blocking=True)
            self.ts[slot].copy_(host_t, non_blocking=True)
            self.ready[slot].record(self.copy_stream)
        self.staged = True

    def step(self, next_host_x=None, next_host_t=None):
        if not self.staged:
            raise L.AlignQError("HostFeeder.step(): nothing staged -- call prefetch(x, t) first")
        cur = self.slot
        main = torch.cuda.current_stream()
        main.wait_event(self.ready[cur])
        self.staged = False
        if next_host_x is not None:                                # overlaps with the step launched just below
            self.prefetch(next_host_x, next_host_t, slot=cur ^ 1)
        loss = self.qat.step(self.xs[cur], self.ts[cur])
        self.consumed[cur].record(main)
        self.slot = cur ^ 1
        return loss
