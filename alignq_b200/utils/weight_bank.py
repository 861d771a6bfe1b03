"""Model-level multi-tensor weight quantization: ONE ``alignq_wq_forward`` launch pair per step for
every quantized conv / linear weight of a model, instead of one pair per layer.

The reference quantizes each layer's weight inside ``Conv2d_Q.forward`` (QA:116-120): 2 reductions and
~25 element-wise ATen kernels per layer per step on tensors of 432 ... 2.4 M elements (launch-bound).
Weights only change in ``optimizer.step()``, so the bank flattens all of them into one fp32 buffer
(the parameters become views of it; names, shapes and ``state_dict`` are unchanged), quantizes the
whole buffer at the start of the forward, and each layer's ``weight_quantize_fn`` picks up its slice
(``weight_q`` / ``weight_cdf`` / ``weight_pdf`` exactly as the per-layer path stores them).  The
backward stays per layer (it needs that layer's upstream gradient).
"""
from __future__ import annotations

import torch

from .. import _lib as L
from ..model.quantization import _single_plan, weight_quantize_fn
from .options import args


class _BankSliceFn(torch.autograd.Function):
    """Identity-forward link between a layer's parameter and its slice of the bank's output."""

    @staticmethod
    def forward(ctx, w, bank, i):
        ctx.bank, ctx.i = bank, i
        ctx.save_for_backward(w)
        ctx.set_materialize_grads(False)
        # a fresh alias: autograd attaches grad_fn to the returned object, and attaching it to the bank's
        # persistent view would keep every iteration's graph (and its AccumulateGrad nodes) alive
        out = bank.wq[i].detach()
        # where a side-stream weight-gradient kernel may deposit d loss / d weight_q of this layer directly (its slice
        # of the bank's flat upstream-gradient buffer), and whom to tell (model/conv_tc.py:_ConvQFn)
        out._alignq_gup = (bank, i)
        return out

    @staticmethod
    def backward(ctx, g):
        if g is None:
            return None, None, None
        (w,) = ctx.saved_tensors
        bank, i = ctx.bank, ctx.i
        g = L.like_layout(g, w, "grad of quantized weight")
        if bank.batched_backward:                 # deposit; WeightBank.flush_backward() runs ONE launch pair
            bank.pending[i] = g
            return None, None, None
        seg_off, chunk_seg, seg_chunk0, nchunks = _single_plan(w.numel(), w.device)
        gw = torch.empty_like(w)
        ws = torch.empty(2 * max(nchunks, 1), dtype=torch.float64, device=w.device)
        with torch.cuda.device_of(w):
            L.check(L.load().alignq_wq_backward(
                w.data_ptr(), g.data_ptr(), 0, seg_off.data_ptr(), chunk_seg.data_ptr(), seg_chunk0.data_ptr(), 1,
                nchunks, bank.w_bit, bank.stats[4 * i: 4 * i + 4].data_ptr(), gw.data_ptr(), 0, ws.data_ptr(),
                L.stream_ptr()), "alignq_wq_backward")
        return gw, None, None


class WeightBank:
    def __init__(self, model):
        self.fns = []
        params = []
        for m in model.modules():
            q = getattr(m, "quantize_fn", None)
            if isinstance(q, weight_quantize_fn) and hasattr(m, "weight") and q.w_bit < 32:
                self.fns.append(q)
                params.append(m.weight)
        if not params:
            raise L.AlignQError("WeightBank: the model has no quantized conv / linear layers")
        bits = {q.w_bit for q in self.fns}
        variants = {q.variant for q in self.fns}
        if len(bits) != 1 or len(variants) != 1:
            raise L.AlignQError("WeightBank needs one bit-width and one variant for all layers")
        self.w_bit, self.variant = bits.pop(), variants.pop()
        self.params = params
        dev = params[0].device
        sizes = [p.numel() for p in params]
        self.flat = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
        off = 0
        for p in params:                                   # re-point every parameter at its slice (same shape,
            n = p.numel()                                  # same strides: NCHW-contiguous or channels_last)
            if not L.is_dense(p.data):
                p.data = p.data.contiguous()
            self.flat[off: off + n].copy_(L.phys(p.data))
            p.data = torch.as_strided(self.flat, p.shape, p.stride(), off)
            off += n
        seg_off, chunk_seg, seg_chunk0, self.nchunks = L.plan_chunks(sizes)
        self.seg_off_host = seg_off
        self.seg_off = torch.tensor(seg_off, dtype=torch.int64, device=dev)
        self.chunk_seg = torch.tensor(chunk_seg, dtype=torch.int32, device=dev)
        self.seg_chunk0 = torch.tensor(seg_chunk0, dtype=torch.int32, device=dev)
        self.wq_flat = torch.empty_like(self.flat)
        self.cdf_flat = torch.empty_like(self.flat)
        self.pdf_flat = torch.empty_like(self.flat)
        self.stats = torch.empty(4 * len(params), dtype=torch.float32, device=dev)
        self.ws = torch.empty(2 * self.nchunks, dtype=torch.float64, device=dev)
        view = lambda flat: [torch.as_strided(flat, p.shape, p.stride(), seg_off[i]) for i, p in enumerate(params)]
        self.wq, self.cdf, self.pdf = view(self.wq_flat), view(self.cdf_flat), view(self.pdf_flat)
        self.fresh = False
        # batched backward: layers deposit their upstream gradients, flush_backward() runs one launch pair
        self.batched_backward = False
        self.pending = [None] * len(params)
        self.gw_flat = torch.zeros_like(self.flat)
        self.gw = view(self.gw_flat)
        self.bwd_ws = torch.empty(2 * self.nchunks, dtype=torch.float64, device=dev)
        # upstream gradients d loss / d weight_q, one flat buffer with the bank's segmentation: the conv layers' weight-
        # gradient kernels write their slices on the side stream, and in data-parallel runs the slices are all-reduced
        # THERE, bucket by bucket while the main backward chain is still running (the quantizer backward is linear in the
        # upstream gradient and the weights are replicated, so reducing before or after it is the same sum)
        self.gup_flat = torch.zeros_like(self.flat)
        self.gup = view(self.gup_flat)
        self.dp_group, self.dp_world = None, 1
        self.buckets, self.bucket_of = [], []
        self.done = [False] * len(params)
        self.reduced = [False] * 0
        for i, q in enumerate(self.fns):
            q._bank = (self, i)

    def enable_dp_overlap(self, group, world, nbuckets=3):
        """Data parallel: all-reduce the upstream weight gradients on the side stream in `nbuckets` contiguous layer
        ranges of roughly equal size, each as soon as all of its layers have deposited."""
        self.dp_group, self.dp_world = group, int(world)
        n = len(self.params)
        total = self.flat.numel()
        self.buckets, self.bucket_of = [], [0] * n
        lo = 0
        for b in range(nbuckets):
            target = total * (b + 1) // nbuckets
            hi = lo
            while hi < n and (self.seg_off_host[hi + 1] <= target or hi == lo):
                hi += 1
            if b == nbuckets - 1:
                hi = n
            if hi > lo:
                self.buckets.append((lo, hi))
                for i in range(lo, hi):
                    self.bucket_of[i] = len(self.buckets) - 1
            lo = hi
        self.reduced = [False] * len(self.buckets)

    def begin_step(self):
        self.done = [False] * len(self.params)
        self.reduced = [False] * len(self.buckets)

    def _reduce_bucket(self, b):
        import torch.distributed as dist
        lo, hi = self.buckets[b]
        a, e = self.seg_off_host[lo], self.seg_off_host[hi]
        dist.all_reduce(self.gup_flat[a:e], op=dist.ReduceOp.SUM, group=self.dp_group)
        self.reduced[b] = True

    def wgrad_deposited(self, i):
        """Called on the SIDE stream right after layer i's weight-gradient kernel was enqueued there."""
        self.done[i] = True
        if self.dp_world > 1 and self.buckets:
            b = self.bucket_of[i]
            lo, hi = self.buckets[b]
            if not self.reduced[b] and all(self.done[lo:hi]):
                self._reduce_bucket(b)

    def finish_dp_reduce(self, side_stream):
        """Before the join: buckets whose layers did not all run through the side stream (e.g. a layer without gradient)
        are reduced now, on the side stream; slices of layers that never deposited are zero."""
        if self.dp_world > 1 and self.buckets:
            with torch.cuda.stream(side_stream):
                for b in range(len(self.buckets)):
                    if not self.reduced[b]:
                        lo, hi = self.buckets[b]
                        for i in range(lo, hi):
                            if not self.done[i]:
                                self.gup[i].zero_()
                        self._reduce_bucket(b)

    def quantize_all(self):
        """One multi-tensor launch pair for every weight of the model (call before the forward)."""
        want = bool(args.store_weight_attrs)
        with torch.cuda.device_of(self.flat):
            L.check(L.load().alignq_wq_forward(
                self.flat.data_ptr(), self.seg_off.data_ptr(), self.chunk_seg.data_ptr(), self.seg_chunk0.data_ptr(),
                len(self.params), self.nchunks, self.w_bit, L.VARIANT_ID[self.variant], self.wq_flat.data_ptr(),
                self.cdf_flat.data_ptr() if want else 0, self.pdf_flat.data_ptr() if want else 0, 0,
                self.stats.data_ptr(), self.ws.data_ptr(), L.stream_ptr()), "alignq_wq_forward (bank)")
        self.fresh = True
        self.begin_step()
        if self.batched_backward:
            self.gw_flat.zero_()                           # flush_backward() always accumulates

    @torch.no_grad()
    def flush_backward(self):
        """After a ``.backward()`` call: weight gradients of every layer that received one, in ONE
        multi-tensor launch pair, accumulated into the bank's gradient buffer (the ``p.grad`` views)."""
        if not self.batched_backward or not any(g is not None for g in self.pending):
            return
        import ctypes
        ptrs = (ctypes.c_void_p * len(self.pending))(*[None if g is None else g.data_ptr() for g in self.pending])
        with torch.cuda.device_of(self.flat):              # pointers travel as kernel arguments: nothing to upload
            L.check(L.load().alignq_wq_backward(
                self.flat.data_ptr(), 0, ptrs, self.seg_off.data_ptr(), self.chunk_seg.data_ptr(),
                self.seg_chunk0.data_ptr(), len(self.params), self.nchunks, self.w_bit, self.stats.data_ptr(),
                self.gw_flat.data_ptr(), 1, self.bwd_ws.data_ptr(), L.stream_ptr()), "alignq_wq_backward (bank)")
        for i, g in enumerate(self.pending):
            if g is not None:
                self.params[i].grad = self.gw[i]
        self._keep = self.pending                          # upstream gradients stay alive until the next flush
        self.pending = [None] * len(self.params)

    def lookup(self, i, w):
        """Called from weight_quantize_fn.forward: the layer's slice, linked into the autograd graph."""
        q = self.fns[i]
        wq = _BankSliceFn.apply(w, self, i)
        q.weight_q = wq.detach()
        q.weight_cdf = self.cdf[i] if args.store_weight_attrs else None
        q.weight_pdf = self.pdf[i] if args.store_weight_attrs else None
        return wq

    def release(self):
        for q in self.fns:
            q._bank = None
