"""Drop-in mirror of the reference's ``model/quantization.py`` (all three variants) on sm_100a kernels.

Same names, constructor arguments, attributes and return conventions as
  QA  cdf_alignment/*/model/quantization.py                              (variant 'A')
  QB  cdf_alignment_admm/resnet-{20,56}-cifar-10/model/quantization.py    (variant 'B')
  QC  cdf_alignment_admm/{dann,dsan}_office/model/quantization.py         (variant 'C')
with every forward/backward executed by the hand-written CUDA kernels behind
``include/alignq_b200.h``.  There is no eager / CPU path: CPU tensors raise ``AlignQError``.

Differences from the reference, all deliberate (DESIGN.md section "Deviations"):
  * config comes from ``alignq_b200.utils.options.args`` (``set_args``), not argparse-at-import;
    the math variant is ``args.variant`` or the ``variant=`` keyword of each constructor;
  * the device is the input tensor's device, not ``cuda:{args.gpus[0]}`` fixed at import (QA:12);
  * ``weight_cdf`` / ``weight_pdf`` / ``weight_q`` are stored in every variant (QA keeps them as
    locals, which breaks cdf_alignment/*/main.py:308-309 as shipped);
  * NaN / std == 0 propagate IEEE-style instead of raising ValueError after a host sync.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib as L
from ..utils.options import args

__all__ = ["uniform_quantize", "cdf", "weight_quantize_fn", "activation_quantize_fn",
           "activation_quantize_fn2", "corr", "conv2d_Q_fn", "linear_Q_fn"]


def _variant(v):
    v = args.variant if v is None else v
    if v not in L.VARIANT_ID:
        raise ValueError(f"variant must be one of {list(L.VARIANT_ID)}, got {v!r}")
    return v


# ------------------------------------------------------------------------------------------------
# uniform_quantize(k)                                                              QA:15-34
# ------------------------------------------------------------------------------------------------
class _UniformQ(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k):
        if k == 32:
            return x.view_as(x)
        xc = L.dev_f32(x, "uniform_quantize input")
        y = torch.empty_like(xc)
        with torch.cuda.device_of(xc):
            L.check(L.load().alignq_uniform_q_fwd(xc.data_ptr(), y.data_ptr(), xc.numel(), k, L.stream_ptr()),
                    "alignq_uniform_q_fwd")
        return y

    @staticmethod
    def backward(ctx, g):
        return g.clone(), None          # straight-through (QA:29-32)


def uniform_quantize(k):
    def qfn(input):
        return _UniformQ.apply(input, k)
    return qfn


# ------------------------------------------------------------------------------------------------
# cdf(m, s, quant_src)                                                     QA:37-50, QB:41-59
# ------------------------------------------------------------------------------------------------
class _CdfFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, m, s, variant, src_is_act, act_range):
        xc = L.dev_f32(x, "cdf input")
        m = L.dev_f32(m.reshape(-1)[:1], "cdf mean")
        s = L.dev_f32(s.reshape(-1)[:1], "cdf std")
        c, p = torch.empty_like(xc), torch.empty_like(xc)
        with torch.cuda.device_of(xc):
            L.check(L.load().alignq_cdf_fwd(xc.data_ptr(), m.data_ptr(), s.data_ptr(), variant, src_is_act,
                                            act_range, c.data_ptr(), p.data_ptr(), xc.numel(), L.stream_ptr()),
                    "alignq_cdf_fwd")
        ctx.save_for_backward(xc, m, s)
        ctx.cfg = (variant, src_is_act, act_range)
        return c, p

    @staticmethod
    def backward(ctx, gc, gp):
        xc, m, s = ctx.saved_tensors
        variant, src_is_act, act_range = ctx.cfg
        gc = None if gc is None else L.dev_f32(gc, "grad cdf")
        gp = None if gp is None else L.dev_f32(gp, "grad pdf")
        gx = torch.empty_like(xc)
        with torch.cuda.device_of(xc):
            L.check(L.load().alignq_cdf_bwd(xc.data_ptr(), m.data_ptr(), s.data_ptr(), variant, src_is_act,
                                            act_range, L.ptr(gc), L.ptr(gp), gx.data_ptr(), xc.numel(),
                                            L.stream_ptr()), "alignq_cdf_bwd")
        return gx, None, None, None, None, None


class cdf(nn.Module):
    """``cdf(m, s, quant_src)(t) -> (mapped_cdf, pdf)``.  Stand-alone form: the gradient flows to
    ``t`` only (m, s are constants); the fused weight path differentiates through mean/std."""

    def __init__(self, m, s, quant_src, variant=None):
        super().__init__()
        self.m, self.s, self.quant_src = m, s, quant_src
        self.variant = _variant(variant)

    def forward(self, tensor):
        for name, v in (("m", self.m), ("s", self.s)):
            if torch.is_tensor(v) and v.requires_grad and torch.is_grad_enabled():
                # the reference's cdf differentiates through m and s (weight_quantize_fn passes mean/std of the
                # weight, QA:70); that chain is implemented by the fused weight_quantize_fn kernels, not here
                raise L.AlignQError(f"cdf(m, s, src): `{name}` requires grad; the stand-alone cdf treats m and s as "
                                    "constants -- use weight_quantize_fn (gradient through mean/std) or detach them")
        m = torch.as_tensor(self.m, dtype=torch.float32, device=tensor.device).detach()
        s = torch.as_tensor(self.s, dtype=torch.float32, device=tensor.device).detach()
        return _CdfFn.apply(tensor, m, s, L.VARIANT_ID[self.variant], int(self.quant_src == "a"),
                            float(args.act_range))


# ------------------------------------------------------------------------------------------------
# weight_quantize_fn(w_bit, stage)                                         QA:52-78, QB:61-85
# ------------------------------------------------------------------------------------------------
_plan_cache = {}


def _single_plan(numel: int, device):
    """Device-resident chunk tables for a one-tensor launch (cached per size and device)."""
    key = (numel, device.index)
    plan = _plan_cache.get(key)
    if plan is None:
        seg_off, chunk_seg, seg_chunk0, nchunks = L.plan_chunks([numel])
        plan = (torch.tensor(seg_off, dtype=torch.int64, device=device),
                torch.tensor(chunk_seg if nchunks else [0], dtype=torch.int32, device=device),
                torch.tensor(seg_chunk0, dtype=torch.int32, device=device), nchunks)
        _plan_cache[key] = plan
    return plan


class _WeightQuantFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w, w_bit, variant, want_attrs):
        wc = L.dev_f32_dense(w, "weight")          # statistics and map are order-independent: any dense layout
        seg_off, chunk_seg, seg_chunk0, nchunks = _single_plan(wc.numel(), wc.device)
        wq = torch.empty_like(wc)
        w_cdf = torch.empty_like(wc) if want_attrs else None
        w_pdf = torch.empty_like(wc) if want_attrs else None
        stats = torch.empty(4, dtype=torch.float32, device=wc.device)
        ws = torch.empty(2 * max(nchunks, 1), dtype=torch.float64, device=wc.device)
        with torch.cuda.device_of(wc):
            L.check(L.load().alignq_wq_forward(
                wc.data_ptr(), seg_off.data_ptr(), chunk_seg.data_ptr(), seg_chunk0.data_ptr(), 1, nchunks,
                w_bit, variant, wq.data_ptr(), L.ptr(w_cdf), L.ptr(w_pdf), 0, stats.data_ptr(), ws.data_ptr(),
                L.stream_ptr()), "alignq_wq_forward")
        ctx.save_for_backward(wc, stats)
        ctx.w_bit = w_bit
        ctx.set_materialize_grads(False)           # no zero-filled grads for the non-differentiable attrs
        outs = (wq, w_cdf, w_pdf) if want_attrs else (wq,)
        if want_attrs:
            ctx.mark_non_differentiable(w_cdf, w_pdf)
        return outs

    @staticmethod
    def backward(ctx, g, *_):
        wc, stats = ctx.saved_tensors
        if g is None:
            return None, None, None, None
        g = L.like_layout(g, wc, "grad of quantized weight")
        seg_off, chunk_seg, seg_chunk0, nchunks = _single_plan(wc.numel(), wc.device)
        gw = torch.empty_like(wc)
        ws = torch.empty(2 * max(nchunks, 1), dtype=torch.float64, device=wc.device)
        with torch.cuda.device_of(wc):
            L.check(L.load().alignq_wq_backward(
                wc.data_ptr(), g.data_ptr(), 0, seg_off.data_ptr(), chunk_seg.data_ptr(), seg_chunk0.data_ptr(), 1,
                nchunks, ctx.w_bit, stats.data_ptr(), gw.data_ptr(), 0, ws.data_ptr(), L.stream_ptr()),
                "alignq_wq_backward")
        return gw, None, None, None


class weight_quantize_fn(nn.Module):
    def __init__(self, w_bit, stage, variant=None):
        super().__init__()
        self.w_bit = w_bit
        self.stage = stage
        self.variant = _variant(variant)
        self.uniform_q = uniform_quantize(k=self.w_bit)
        self._bank = None                            # (WeightBank, index) when a model-level bank owns this weight

    def forward(self, x):
        if self._bank is not None and self._bank[0].fresh and x is self._bank[0].params[self._bank[1]]:
            return self._bank[0].lookup(self._bank[1], x)     # quantized by the bank's multi-tensor launch
        if self.w_bit == 32:                         # QB:73-76
            self.weight_cdf = x
            self.weight_q = x
            return x
        want = bool(args.store_weight_attrs)
        outs = _WeightQuantFn.apply(x, self.w_bit, L.VARIANT_ID[self.variant], want)
        # The attributes are detached copies of the handles: holding the graph-attached output across
        # iterations would keep the parameter's AccumulateGrad node (and the stream it was created on)
        # alive, which breaks CUDA-graph capture after a side-stream warm-up.
        self.weight_q = outs[0].detach()
        self.weight_cdf, self.weight_pdf = (outs[1], outs[2]) if want else (None, None)
        return outs[0]


# ------------------------------------------------------------------------------------------------
# activation_quantize_fn / activation_quantize_fn2            QA:81-103, QB:88-132, QC:87-156
# ------------------------------------------------------------------------------------------------
class _ActQuantFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, a_bit, act_range, variant, return_cdf):
        xc = L.dev_f32_dense(x, "activation")
        y = torch.empty_like(xc)
        with torch.cuda.device_of(xc):
            L.check(L.load().alignq_act_fwd(xc.data_ptr(), y.data_ptr(), 0, xc.numel(), a_bit, act_range,
                                            variant, return_cdf, L.stream_ptr()), "alignq_act_fwd")
        ctx.save_for_backward(xc)
        ctx.cfg = (a_bit, act_range, variant, return_cdf)
        return y

    @staticmethod
    def backward(ctx, gy):
        (xc,) = ctx.saved_tensors
        a_bit, act_range, variant, return_cdf = ctx.cfg
        gy = L.like_layout(gy, xc, "grad of quantized activation")
        gx = torch.empty_like(xc)
        with torch.cuda.device_of(xc):
            L.check(L.load().alignq_act_bwd(xc.data_ptr(), gy.data_ptr(), gx.data_ptr(), xc.numel(), a_bit,
                                            act_range, variant, return_cdf, L.stream_ptr()), "alignq_act_bwd")
        return gx, None, None, None, None


_ws_cache = {}


def _gram_ws(B: int, Fdim: int, device) -> torch.Tensor:
    """One shared scratch buffer per device, grown on demand (stream-ordered reuse across layers)."""
    need = int(L.load().alignq_gram_ws_bytes(B, Fdim))
    ws = _ws_cache.get(device.index)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=device)
        _ws_cache[device.index] = ws
    return ws


class _ActAdmmFn(torch.autograd.Function):
    """y, trans_loss, D = fused activation quantizer + corr(x) / corr(t) + ADMM loss."""

    @staticmethod
    def forward(ctx, x, alterD, gamma, a_bit, act_range, eps, mu, rho, gram_mode, variant_id=1, d_out=None,
                param_grads=True):
        xc = L.dev_f32_dense(x, "activation")      # per-sample feature order is irrelevant to the Gram
        Z = L.dev_f32(alterD, "alterD")
        U = L.dev_f32(gamma, "gamma")
        B = xc.shape[0]
        Fdim = xc.numel() // B
        dim = Z.shape[0]
        if Z.shape != (dim, dim) or U.shape != (dim, dim) or dim < B:
            raise L.AlignQError(f"ADMM dim {tuple(Z.shape)} must be square and >= batch {B}")
        y = torch.empty_like(xc)
        D = d_out if (d_out is not None and d_out.shape == (B, B) and d_out.is_contiguous()) \
            else torch.empty(B, B, dtype=torch.float32, device=xc.device)
        dLdD = torch.empty_like(D)
        loss = torch.empty((), dtype=torch.float32, device=xc.device)
        ws = _gram_ws(B, Fdim, xc.device)
        with torch.cuda.device_of(xc):
            L.check(L.load().alignq_act_admm_fwd(
                xc.data_ptr(), B, Fdim, a_bit, act_range, eps, Z.data_ptr(), U.data_ptr(), dim, mu, rho,
                y.data_ptr(), D.data_ptr(), loss.data_ptr(), dLdD.data_ptr(), ws.data_ptr(), ws.numel(),
                gram_mode, L.stream_ptr()), "alignq_act_admm_fwd")
        ctx.save_for_backward(xc, dLdD, D, Z, U)
        ctx.cfg = (a_bit, act_range, eps, mu, rho, gram_mode, variant_id, bool(param_grads))
        ctx.mark_non_differentiable(D)
        ctx.set_materialize_grads(False)       # a backward pass that does not reach trans_loss skips the Gram backward
        return y, loss, D

    @staticmethod
    def backward(ctx, gy, gloss, _gD):
        xc, dLdD, D, Z, U = ctx.saved_tensors
        a_bit, act_range, eps, mu, rho, gram_mode, variant_id, param_grads = ctx.cfg
        B = xc.shape[0]
        Fdim = xc.numel() // B
        dim = Z.shape[0]
        lib = L.load()
        gx = gZ = gU = None
        with torch.cuda.device_of(xc):
            if gloss is None:
                # this backward pass does not involve trans_loss (e.g. CE.backward(retain_graph=True) in
                # cdf_alignment_admm/.../main.py:300): only the straight-through quantizer backward remains
                if gy is not None and ctx.needs_input_grad[0]:
                    gyc = L.like_layout(gy, xc, "grad of quantized activation")
                    gx = torch.empty_like(xc)
                    L.check(lib.alignq_act_bwd(xc.data_ptr(), gyc.data_ptr(), gx.data_ptr(), xc.numel(), a_bit,
                                               act_range, variant_id, 0, L.stream_ptr()), "alignq_act_bwd")
                return (gx,) + (None,) * 11
            gl = L.dev_f32(gloss.reshape(1), "grad of trans_loss")
            if ctx.needs_input_grad[0]:
                gyc = None if gy is None else L.like_layout(gy, xc, "grad of quantized activation")
                gx = torch.empty_like(xc)
                ws = _gram_ws(B, Fdim, xc.device)
                L.check(lib.alignq_act_admm_bwd(
                    xc.data_ptr(), L.ptr(gyc), dLdD.data_ptr(), gl.data_ptr(), B, Fdim, a_bit, act_range, eps,
                    gx.data_ptr(), ws.data_ptr(), ws.numel(), gram_mode, L.stream_ptr()), "alignq_act_admm_bwd")
            want_z = ctx.needs_input_grad[1] and param_grads
            want_u = ctx.needs_input_grad[2] and param_grads
            if want_z or want_u:
                gZ = torch.empty_like(Z) if want_z else None
                gU = torch.empty_like(U) if want_u else None
                L.check(lib.alignq_admm_loss(D.data_ptr(), B, Z.data_ptr(), U.data_ptr(), dim, 1, mu, rho,
                                             gl.data_ptr(), 0, 0, 0, L.ptr(gZ), L.ptr(gU), L.stream_ptr()),
                        "alignq_admm_loss (parameter grads)")
        return (gx, gZ, gU) + (None,) * 9


class activation_quantize_fn(nn.Module):
    """QA / QC form: ``activation_quantize_fn(a_bit, stage)(x) -> y``.
    QB form: ``activation_quantize_fn(a_bit, stage, admm)(x) -> (y, trans_loss)``."""

    _returns_tuple_with_admm = True

    def __init__(self, a_bit, stage, admm=None, variant=None):
        super().__init__()
        self.a_bit = a_bit
        self.stage = stage
        self.variant = _variant(variant)
        if admm is not None and self.variant == "A":
            self.variant = "B"            # an ADMM term implies the symmetric map of QB/QC
        self.uniform_q = uniform_quantize(k=a_bit)
        self.opt = admm

    def _plain(self, x):
        return _ActQuantFn.apply(x, self.a_bit, float(args.act_range), L.VARIANT_ID[self.variant],
                                 int(self.a_bit == 32))

    def forward(self, x):
        tuple_out = self.opt is not None
        if self.a_bit == 32 and self.stage != "align":           # QA:92-95 / QB:103-107
            return (x, 0) if tuple_out else x
        if tuple_out and args.method == "ours" and self.a_bit < 32:   # QB:112-123
            eps = 0.0 if self.variant == "B" else 1e-5
            if args.dp_gram == "feature":
                from ..utils import dp_gram
                if dp_gram.world() > 1:                # global-batch Gram over the ranks' feature slices
                    return dp_gram.feature_sharded_act_admm(x, self.opt, self.a_bit, float(args.act_range), eps,
                                                            L.GRAM_MODE_ID[args.gram_mode], L.VARIANT_ID[self.variant])
            y, loss, D = _ActAdmmFn.apply(x, self.opt.alterD, self.opt.gamma, self.a_bit, float(args.act_range),
                                          eps, float(self.opt.mu), float(self.opt.rho),
                                          L.GRAM_MODE_ID[args.gram_mode], L.VARIANT_ID[self.variant],
                                          self.opt.d_slot(x.shape[0]) if hasattr(self.opt, "d_slot") else None,
                                          bool(getattr(self.opt, "param_grads", True)) and bool(args.admm_param_grads))
            self.opt.D = D
            return y, loss
        y = self._plain(x)
        return (y, 0) if tuple_out else y


class activation_quantize_fn2(activation_quantize_fn):
    """QC's ADMM-enabled activation quantizer (QC:112-156): always returns ``(y, trans_loss)``."""

    def __init__(self, a_bit, stage, admm, variant=None):
        super().__init__(a_bit, stage, admm, variant if variant is not None else
                         ("C" if args.variant == "A" else args.variant))


# ------------------------------------------------------------------------------------------------
# corr(x, y)                                                          QB:134-137, QC:158-161
# ------------------------------------------------------------------------------------------------
class _CorrFn(torch.autograd.Function):
    """corr(x, y) with the reference's autograd semantics: differentiable through mean and std of both
    operands (QB:135-137).  Backward: ``alignq_corr_bwd`` (fp32 FFMA kernels)."""

    @staticmethod
    def forward(ctx, x, y, eps, gram_mode, same):
        xc = L.dev_f32(x, "corr x")
        yc = xc if same else L.dev_f32(y, "corr y")
        B, Fdim = xc.shape
        G = torch.empty(B, B, dtype=torch.float32, device=xc.device)
        ws = _gram_ws(B, Fdim, xc.device)
        with torch.cuda.device_of(xc):
            L.check(L.load().alignq_corr_fwd(xc.data_ptr(), yc.data_ptr(), B, Fdim, float(eps), G.data_ptr(),
                                             ws.data_ptr(), ws.numel(), gram_mode, L.stream_ptr()), "alignq_corr_fwd")
        ctx.save_for_backward(xc, yc)
        ctx.cfg = (float(eps), bool(same))
        return G

    @staticmethod
    def backward(ctx, dG):
        xc, yc = ctx.saved_tensors
        eps, same = ctx.cfg
        B, Fdim = xc.shape
        dG = L.dev_f32(dG, "grad of corr")
        want_x = ctx.needs_input_grad[0] or (same and ctx.needs_input_grad[1])
        want_y = (not same) and ctx.needs_input_grad[1]
        gx = torch.empty_like(xc) if want_x else None
        gy = torch.empty_like(yc) if want_y else None
        if B < 2:
            raise L.AlignQError("corr backward needs a batch of at least 2 rows (unbiased std)")
        ws = _gram_ws(B, Fdim, xc.device)
        with torch.cuda.device_of(xc):
            L.check(L.load().alignq_corr_bwd(xc.data_ptr(), yc.data_ptr(), dG.data_ptr(), B, Fdim, eps, L.ptr(gx),
                                             L.ptr(gy), ws.data_ptr(), ws.numel(), L.stream_ptr()), "alignq_corr_bwd")
        if same:          # one operand passed twice: autograd adds both slots, so hand the whole gradient to slot 0
            return gx, None, None, None, None
        return gx, gy, None, None, None


def corr(x, y, eps=None):
    """[B,F],[B,F] -> [B,B].  eps defaults to the variant's (0 for 'A'/'B', 1e-5 for 'C').  Differentiable with
    respect to both operands like the reference's (QB:134-137); the training path uses the fused
    activation_quantize_fn instead, which never materialises the standardised operands."""
    if eps is None:
        eps = 1e-5 if args.variant == "C" else 0.0
    if x.dim() != 2 or x.shape != y.shape:
        raise L.AlignQError(f"corr expects two [B, F] matrices of equal shape, got {tuple(x.shape)}, {tuple(y.shape)}")
    return _CorrFn.apply(x, y, float(eps), L.GRAM_MODE_ID[args.gram_mode], y is x)


# ------------------------------------------------------------------------------------------------
# conv2d_Q_fn / linear_Q_fn           QA:107-122; cdf_alignment/dann_office/model/resnet.py:148-160
# ------------------------------------------------------------------------------------------------
def conv2d_Q_fn(w_bit, stage, variant=None):
    class Conv2d_Q(nn.Conv2d):
        def __init__(self, in_channels, out_channels, kernel_size, stride=1,
                     padding=0, dilation=1, groups=1, bias=True):
            super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
            self.quantize_fn = weight_quantize_fn(w_bit=w_bit, stage=stage, variant=variant)

        def forward(self, input, order=None):
            weight_q = self.quantize_fn(self.weight)
            if args.async_wgrad and self.bias is None and input.is_cuda and self.padding_mode == "zeros":
                from . import conv_tc                                 # weight gradient on a side stream (QATStep joins)
                return conv_tc.conv_async_wgrad(input, weight_q, self.stride, self.padding, self.dilation, self.groups)
            if args.own_conv != "off":
                from . import conv_tc
                if conv_tc.applies(input, weight_q, self.stride, self.padding, self.dilation, self.groups, self.bias):
                    return conv_tc.conv3x3(input, weight_q)           # hand-written tcgen05 kernels
                if conv_tc.applies_stem(input, weight_q, self.stride, self.padding, self.dilation, self.groups, self.bias):
                    return conv_tc.stem_conv(input, weight_q)          # first layer: direct fp32 kernels
            return F.conv2d(input, weight_q, self.bias, self.stride, self.padding, self.dilation, self.groups)

    return Conv2d_Q


def linear_Q_fn(w_bit, stage, variant=None):
    class Linear_Q(nn.Linear):
        def __init__(self, in_features, out_features, bias=True):
            super().__init__(in_features, out_features, bias)
            self.w_bit = w_bit
            self.quantize_fn = weight_quantize_fn(w_bit=w_bit, stage=stage, variant=variant)

        def forward(self, input):
            weight_q = self.quantize_fn(self.weight)
            return F.linear(input, weight_q, self.bias)

    return Linear_Q
