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
        return bank.wq[i].detach()

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
        for i, q in enumerate(self.fns):
            q._bank = (self, i)

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
