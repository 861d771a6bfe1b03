"""``F.conv2d(input, weight_q, None, 1, 1)`` of ``Conv2d_Q.forward`` (QA:116-120) on the hand-written tcgen05 kernels
(csrc/conv_tc.cu): forward, data gradient and weight gradient of the 3x3 / stride 1 / padding 1 / Cin == Cout in
{16, 32, 64} convolutions on channels_last fp32 tensors -- SURVEY.md 8(f) item 2.  Every other convolution of the model
files (stem, strided, 1x1 projection, depthwise, NCHW inputs) keeps calling the library convolution, as the reference does.

``args.own_conv``: "off" (library convolution everywhere), "tf32" (one tensor-core pass on tf32 operands: the numerics of
cuDNN under torch's default ``allow_tf32 = True``) or "tf32x3" (three passes on H + L split operands: fp32 parity, 1e-5).
``own_conv_channels`` routes forward + both gradients, ``own_dgrad_channels`` / ``own_wgrad_channels`` one gradient alone.

Besides the convolution itself this module carries the step-level plumbing around it: weight gradients on side streams
(``WgradStream``), the BatchNorm statistics of the following layer from the forward epilogue (``conv_with_bn_stats``), and
-- ``args.fuse_dgrad_bn`` -- the backward reduce pass of the PRECEDING fused bn-act layer from the data-gradient epilogue:
``_ConvQFn.backward`` finds that layer's saved tensors in the ``link`` its output carries (model/fused.py), adds a parked
shortcut gradient, and leaves the affine gradients and the two means for ``alignq_bn_act_bwd_apply``.
"""
from __future__ import annotations

import os

import torch

from .. import _lib as L
from ..utils.options import args

_ws = {}


def _workspace(C, device):
    key = (C, device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _ws.get(key)
    if ws is None:
        ws = torch.empty(int(L.load().alignq_conv3x3_ws_bytes(C)), dtype=torch.uint8, device=device)
        _ws[key] = ws
    return ws


def _covered(x, weight, stride, padding, dilation, groups, bias, channels) -> bool:
    if args.own_conv == "off" or bias is not None or groups != 1:
        return False
    if tuple(stride) != (1, 1) or tuple(padding) != (1, 1) or tuple(dilation) != (1, 1):
        return False
    if weight.dim() != 4 or tuple(weight.shape[2:]) != (3, 3) or weight.shape[0] != weight.shape[1]:
        return False
    C = weight.shape[0]
    if C not in (16, 32, 64) or C not in tuple(channels) or (C == 64 and args.own_conv == "tf32x3"):
        return False
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == C
            and x.is_contiguous(memory_format=torch.channels_last) and x.data_ptr() % 16 == 0
            and weight.dtype == torch.float32 and weight.data_ptr() % 16 == 0)


def applies(x, weight, stride, padding, dilation, groups, bias) -> bool:
    """Forward + data gradient (+ weight gradient) on the own kernels."""
    return _covered(x, weight, stride, padding, dilation, groups, bias, args.own_conv_channels)


def applies_dgrad(x, weight, stride, padding, dilation, groups, bias) -> bool:
    """The DATA gradient alone on the own kernel (``own_dgrad_channels``): the library's data-gradient kernels for these
    shapes are cluster launches that come with a parameter-copy node in front (2.5 us on the main chain, each)."""
    return (_covered(x, weight, stride, padding, dilation, groups, bias, tuple(args.own_conv_channels) + tuple(args.own_dgrad_channels))
            and weight.is_contiguous(memory_format=torch.channels_last))


def applies_wgrad(x, weight, stride, padding, dilation, groups, bias) -> bool:
    """The WEIGHT gradient alone on the own kernel (``own_wgrad_channels``): at C = 32 / 64 the library's forward and data
    gradient are still the faster ones, its split-K weight gradient (14-21 us per layer, a memset in front) is not."""
    return (_covered(x, weight, stride, padding, dilation, groups, bias, tuple(args.own_conv_channels) + tuple(args.own_wgrad_channels))
            and weight.is_contiguous(memory_format=torch.channels_last))


class _Conv3x3Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, mode):
        N, C, H, W = x.shape
        wc = w if w.is_contiguous(memory_format=torch.channels_last) else w.contiguous(memory_format=torch.channels_last)
        y = torch.empty_like(x)                       # channels_last like x
        with torch.cuda.device_of(x):
            L.check(L.load().alignq_conv3x3_fwd(x.data_ptr(), wc.data_ptr(), y.data_ptr(), N, H, W, C, mode, L.stream_ptr()),
                    "alignq_conv3x3_fwd")
        ctx.save_for_backward(x, wc)
        ctx.mode = mode
        ctx.w_like = w
        return y

    @staticmethod
    def backward(ctx, gy):
        x, wc = ctx.saved_tensors
        N, C, H, W = x.shape
        gy = L.like_layout(gy, x, "grad of conv output")
        lib = L.load()
        gx = gw = None
        with torch.cuda.device_of(x):
            if ctx.needs_input_grad[0]:
                gx = torch.empty_like(x)
                L.check(lib.alignq_conv3x3_bwd_data(gy.data_ptr(), wc.data_ptr(), gx.data_ptr(), N, H, W, C, ctx.mode,
                                                    L.stream_ptr()), "alignq_conv3x3_bwd_data")
            if ctx.needs_input_grad[1]:
                gw = torch.empty_like(wc)             # channels_last [Cout, Cin, 3, 3] = physical [Cout][3][3][Cin]
                ws = _workspace(C, x.device)
                L.check(lib.alignq_conv3x3_bwd_weight(x.data_ptr(), gy.data_ptr(), gw.data_ptr(), N, H, W, C, ctx.mode, 0,
                                                      ws.data_ptr(), ws.numel(), L.stream_ptr()), "alignq_conv3x3_bwd_weight")
                if gw.stride() != ctx.w_like.stride():
                    gw = torch.empty_like(ctx.w_like).copy_(gw)
        return gx, gw, None


def conv3x3(x, weight):
    """3x3 / stride 1 / padding 1 convolution on the tcgen05 kernels; call ``applies`` first."""
    return _Conv3x3Fn.apply(x, weight, L.CONV_MODE_ID[args.own_conv])


# ------------------------------------------------------------------------------------------------
# Weight gradients off the critical path.  In the backward pass of a conv layer only the DATA gradient feeds the next
# layer; the WEIGHT gradient is not needed before the optimizer runs.  With ``args.async_wgrad`` (switched on by
# QATStep together with the weight bank's batched backward, which is the only consumer of these gradients and runs
# after ``join()``) every quantized convolution computes its weight gradient on a side stream -- own tcgen05 kernel or
# the library's -- so the ~40 small, latency-bound weight-gradient kernels of a step overlap the main chain of
# BatchNorm / quantizer / data-gradient kernels instead of sitting in it.  Fork and join are plain CUDA events, so the
# whole pattern is captured into the step's CUDA graph as parallel branches.
class WgradStream:
    _inst = {}
    NSTREAMS = 3          # round-robin: the weight-gradient kernels of neighbouring layers also overlap EACH OTHER, which
                          # shortens the tail left when the main backward chain ends (the largest ones are issued last)

    def __init__(self, device):
        self.sides = [torch.cuda.Stream(device=device) for _ in range(self.NSTREAMS)]
        self.side = self.sides[0]                     # the stream data-parallel bucket reductions are issued on
        self.next = 0
        self.keep = []                                # tensors the side streams still read
        self.used = set()                             # indices of the side streams forked since the last join
        self.dirty = False

    @classmethod
    def get(cls, device):
        key = device.index
        if key not in cls._inst:
            cls._inst[key] = cls(device)
        return cls._inst[key]

    def fork(self, *tensors, single=False):
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        k = 0 if single else self.next % self.NSTREAMS
        side = self.sides[k]
        self.used.add(k)
        self.next += 1
        side.wait_event(ev)
        self.keep.extend(tensors)
        self.dirty = True
        return side

    def join(self):
        """Make the current stream wait for every weight gradient issued so far (call before they are consumed)."""
        if self.dirty:
            main = torch.cuda.current_stream()
            # only the streams that were forked: inside a CUDA-graph capture, waiting for a stream that never joined the
            # capture (data-parallel overlap mode uses side stream 0 alone) is cudaErrorStreamCaptureIsolation
            for k in sorted(self.used):
                main.wait_stream(self.sides[k])
            self.used.clear()
            self.keep.clear()
            self.dirty = False
            self.next = 0


def join_wgrads(bank=None):
    for inst in list(WgradStream._inst.values()):
        if bank is not None and inst.dirty:
            bank.finish_dp_reduce(inst.side)
        inst.join()


class _ConvQFn(torch.autograd.Function):
    """Any bias-free Conv2d_Q convolution with the weight gradient on the side stream."""

    @staticmethod
    def forward(ctx, x, w, stride, padding, dilation, groups, own, mode, bn=None, bn_ws=None, sync=True, up=None, own_w=None,
                own_d=None):
        mean = invstd = None
        own_w = own if own_w is None else bool(own_w)      # the weight / data gradient may take the own kernels on their own
        own_d = own if own_d is None else bool(own_d)
        ctx.up = up                                   # link of the fused bn-act layer that produced x (fused.py), or None
        if own:
            N, C, H, W = x.shape
            wc = w if w.is_contiguous(memory_format=torch.channels_last) else w.contiguous(memory_format=torch.channels_last)
            y = torch.empty_like(x)
            with torch.cuda.device_of(x):
                if bn is None:
                    L.check(L.load().alignq_conv3x3_fwd(x.data_ptr(), wc.data_ptr(), y.data_ptr(), N, H, W, C, mode,
                                                        L.stream_ptr()), "alignq_conv3x3_fwd")
                else:             # batch statistics of the following BatchNorm from the convolution's epilogue
                    mean = torch.empty(C, dtype=torch.float32, device=x.device)
                    invstd = torch.empty(C, dtype=torch.float32, device=x.device)
                    ws, counter = bn_ws
                    L.check(L.load().alignq_conv3x3_fwd_bnstats(
                        x.data_ptr(), wc.data_ptr(), y.data_ptr(), N, H, W, C, mode, L.ptr(bn.running_mean),
                        L.ptr(bn.running_var), float(bn.momentum), float(bn.eps), mean.data_ptr(), invstd.data_ptr(),
                        ws.data_ptr(), counter.data_ptr(), L.ptr(bn.num_batches_tracked), L.stream_ptr()),
                        "alignq_conv3x3_fwd_bnstats")
        else:
            wc = w
            y = torch.ops.aten.convolution(x, w, None, stride, padding, dilation, False, (0, 0), groups)
        ctx.save_for_backward(x, wc)
        ctx.cfg = (tuple(stride), tuple(padding), tuple(dilation), groups, own, mode, bool(sync), own_w, own_d)
        ctx.w_like = w
        ctx.gup = getattr(w, "_alignq_gup", None) if sync else None    # (WeightBank, layer index) or None
        ctx.set_materialize_grads(False)              # no zero-fill launches for the (non-differentiable) statistics outputs
        if bn is None:
            return y
        ctx.mark_non_differentiable(mean, invstd)
        return y, mean, invstd

    @staticmethod
    def backward(ctx, gy, *_unused):
        if gy is None:
            return (None,) * 14
        x, wc = ctx.saved_tensors
        stride, padding, dilation, groups, own, mode, on_side, own_w, own_d = ctx.cfg
        lib = L.load()
        gx = gw = None
        if own_d or own_w:
            gy = L.like_layout(gy, x, "grad of conv output")
        else:
            gy = gy if gy.is_contiguous(memory_format=torch.channels_last) or gy.is_contiguous() else gy.contiguous()
        # Order of the two launches: the weight gradient goes to a side stream either way.  Forked BEFORE the data
        # gradient (default) it runs beside it; ALIGNQ_WGRAD_FIRST=0 forks it after, so that it starts when the data
        # gradient completes (measured on the ResNet-20 step: no difference, 1.079 vs 1.082 ms).
        wgrad_first = os.environ.get("ALIGNQ_WGRAD_FIRST", "1") == "1"

        def weight_gradient():
            nonlocal gw
            if ctx.needs_input_grad[1]:                   # weight gradient first: it runs beside everything that follows
                ws_ = WgradStream.get(x.device)
                # (data-parallel bucket reductions are issued behind the deposits: those layers all use side stream 0)
                dp = ctx.gup is not None and ctx.gup[0].dp_world > 1 and bool(ctx.gup[0].buckets)
                side = ws_.fork(x, gy, wc, single=dp) if on_side else torch.cuda.current_stream()
                slot = None
                if ctx.gup is not None:                   # the bank's flat upstream-gradient buffer: this layer's slice
                    slot = ctx.gup[0].gup[ctx.gup[1]]
                    if slot.shape != wc.shape or slot.stride() != wc.stride() or slot.data_ptr() % 16:
                        slot = None
                with torch.cuda.stream(side):
                    if own_w:
                        N, C, H, W = x.shape
                        gw = slot if slot is not None else torch.empty_like(wc)
                        wsp = _workspace(C, x.device)     # keyed by the side stream: successive launches there are ordered
                        L.check(lib.alignq_conv3x3_bwd_weight(x.data_ptr(), gy.data_ptr(), gw.data_ptr(), N, H, W, C, mode, 0,
                                                              wsp.data_ptr(), wsp.numel(), L.stream_ptr()), "alignq_conv3x3_bwd_weight")
                        if gw.stride() != ctx.w_like.stride():
                            gw = torch.empty_like(ctx.w_like).copy_(gw)
                    else:
                        _, gw, _ = torch.ops.aten.convolution_backward(gy, x, wc, None, stride, padding, dilation, False, (0, 0),
                                                                       groups, (False, True, False))
                        if slot is not None:
                            gw = slot.copy_(gw)
                        elif gw.stride() != ctx.w_like.stride():         # any layout fix-up belongs on the side stream too
                            gw = torch.empty_like(ctx.w_like).copy_(gw)
                    if slot is not None:                  # data parallel: this bucket's all-reduce may start right here
                        ctx.gup[0].wgrad_deposited(ctx.gup[1])
                if on_side:
                    ws_.keep.append(gw)

        def data_gradient():
            nonlocal gx
            if ctx.needs_input_grad[0]:
                if own_d:
                    N, C, H, W = x.shape
                    gx = torch.empty_like(x)
                    fwd = ctx.up.get("fwd") if ctx.up is not None else None
                    if (fwd is not None and C == 16 and fwd[0].shape == x.shape and fwd[0].stride() == x.stride()
                            and fwd[0].dtype == torch.float32 and fwd[0].data_ptr() % 16 == 0 and "reduced" not in ctx.up):
                        # x is the output of a fused bn-act layer and gx its upstream gradient: that layer's backward reduce
                        # pass (and the sum with a parked second gradient) runs in this kernel's epilogue
                        from .fused import _bn_ws, _take_extra
                        bx, mean, invstd, bw, bb, bn, (rows, Cb, a_bit, act_range, variant, relu) = fwd
                        gy2 = _take_extra(ctx.up, x)
                        gw_bn = torch.empty(C, dtype=torch.float32, device=x.device) if bw is not None else None
                        gb_bn = torch.empty(C, dtype=torch.float32, device=x.device) if bb is not None else None
                        ws, counter = _bn_ws(bn, C, x.device)
                        with torch.cuda.device_of(x):
                            L.check(lib.alignq_conv3x3_bwd_data_bnreduce(
                                gy.data_ptr(), wc.data_ptr(), gx.data_ptr(), N, H, W, C, mode, bx.data_ptr(), x.data_ptr(),
                                L.ptr(gy2), mean.data_ptr(), invstd.data_ptr(), L.ptr(bw), L.ptr(bb), float(act_range),
                                int(relu), L.ptr(gw_bn), L.ptr(gb_bn), ws.data_ptr(), counter.data_ptr(), L.stream_ptr()),
                                "alignq_conv3x3_bwd_data_bnreduce")
                        ctx.up["reduced"] = (gx.data_ptr(), gw_bn, gb_bn)
                    else:
                        with torch.cuda.device_of(x):
                            L.check(lib.alignq_conv3x3_bwd_data(gy.data_ptr(), wc.data_ptr(), gx.data_ptr(), N, H, W, C, mode,
                                                                L.stream_ptr()), "alignq_conv3x3_bwd_data")
                else:
                    gx, _, _ = torch.ops.aten.convolution_backward(gy, x, wc, None, stride, padding, dilation, False, (0, 0),
                                                                   groups, (True, False, False))

        for step in ((weight_gradient, data_gradient) if wgrad_first else (data_gradient, weight_gradient)):
            step()
        return gx, gw, None, None, None, None, None, None, None, None, None, None, None, None


_stem_ws = {}


def _stem_workspace(Cout, device):
    """fp64 accumulator copies + ticket of the stem weight gradient: zero before first use, re-armed by the kernel."""
    key = (Cout, device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _stem_ws.get(key)
    if ws is None:
        ws = torch.zeros(int(L.load().alignq_conv3x3_stem_ws_bytes(Cout)), dtype=torch.uint8, device=device)
        _stem_ws[key] = ws
    return ws


def applies_stem(x, weight, stride, padding, dilation, groups, bias) -> bool:
    """First-layer convolution (image, Cin = 3 -> 16 / 32 channels, 3x3, stride 1, padding 1) on csrc/conv_stem.cu."""
    if args.own_conv == "off" or not args.own_conv_stem or bias is not None or groups != 1:
        return False
    if tuple(stride) != (1, 1) or tuple(padding) != (1, 1) or tuple(dilation) != (1, 1):
        return False
    if weight.dim() != 4 or tuple(weight.shape[1:]) != (3, 3, 3) or weight.shape[0] not in (16, 32):
        return False
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 3 and not x.requires_grad
            and x.is_contiguous(memory_format=torch.channels_last) and weight.dtype == torch.float32)


class _StemConvFn(torch.autograd.Function):
    """conv0: direct fp32 forward (optionally with the following BatchNorm's batch statistics) and weight gradient."""

    @staticmethod
    def forward(ctx, x, w, bn, bn_ws, on_side):
        N, _, H, W = x.shape
        Cout = w.shape[0]
        wc = w if w.is_contiguous(memory_format=torch.channels_last) else w.contiguous(memory_format=torch.channels_last)
        y = torch.empty((N, Cout, H, W), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
        mean = invstd = None
        with torch.cuda.device_of(x):
            if bn is None:
                L.check(L.load().alignq_conv3x3_stem_fwd(x.data_ptr(), wc.data_ptr(), y.data_ptr(), N, H, W, Cout, 0, 0, 0.0, 0.0,
                                                         0, 0, 0, 0, 0, L.stream_ptr()), "alignq_conv3x3_stem_fwd")
            else:
                mean = torch.empty(Cout, dtype=torch.float32, device=x.device)
                invstd = torch.empty(Cout, dtype=torch.float32, device=x.device)
                ws, counter = bn_ws
                L.check(L.load().alignq_conv3x3_stem_fwd(
                    x.data_ptr(), wc.data_ptr(), y.data_ptr(), N, H, W, Cout, L.ptr(bn.running_mean), L.ptr(bn.running_var),
                    float(bn.momentum), float(bn.eps), mean.data_ptr(), invstd.data_ptr(), ws.data_ptr(), counter.data_ptr(),
                    L.ptr(bn.num_batches_tracked), L.stream_ptr()), "alignq_conv3x3_stem_fwd")
        ctx.save_for_backward(x, wc)
        ctx.w_like = w
        ctx.on_side = bool(on_side)
        ctx.gup = getattr(w, "_alignq_gup", None) if on_side else None
        ctx.set_materialize_grads(False)
        if bn is None:
            return y
        ctx.mark_non_differentiable(mean, invstd)
        return y, mean, invstd

    @staticmethod
    def backward(ctx, gy, *_unused):
        if gy is None or not ctx.needs_input_grad[1]:
            return (None,) * 5
        x, wc = ctx.saved_tensors
        N, _, H, W = x.shape
        Cout = wc.shape[0]
        if not gy.is_contiguous(memory_format=torch.channels_last):
            gy = gy.contiguous(memory_format=torch.channels_last)
        ws_ = WgradStream.get(x.device)
        dp = ctx.gup is not None and ctx.gup[0].dp_world > 1 and bool(ctx.gup[0].buckets)
        side = ws_.fork(x, gy, wc, single=dp) if ctx.on_side else torch.cuda.current_stream()
        slot = None
        if ctx.gup is not None:
            slot = ctx.gup[0].gup[ctx.gup[1]]
            if slot.shape != wc.shape or slot.stride() != wc.stride():
                slot = None
        with torch.cuda.stream(side):
            gw = slot if slot is not None else torch.empty_like(wc)
            wsp = _stem_workspace(Cout, x.device)
            L.check(L.load().alignq_conv3x3_stem_bwd_weight(x.data_ptr(), gy.data_ptr(), gw.data_ptr(), N, H, W, Cout,
                                                            wsp.data_ptr(), wsp.numel(), L.stream_ptr()),
                    "alignq_conv3x3_stem_bwd_weight")
            if gw.stride() != ctx.w_like.stride():
                gw = torch.empty_like(ctx.w_like).copy_(gw)
            if slot is not None:
                ctx.gup[0].wgrad_deposited(ctx.gup[1])
        if ctx.on_side:
            ws_.keep.append(gw)
        return None, gw, None, None, None


def stem_conv(x, weight, bn=None, bn_ws=None):
    """conv0 on the own kernels; with ``bn`` (training mode) also returns that BatchNorm's (save_mean, save_invstd)."""
    return _StemConvFn.apply(x, weight, bn, bn_ws, bool(args.async_wgrad))


def _up_link(x):
    """The link of the fused bn-act layer whose output x is (model/fused.py), when its backward reduce pass may move into
    this convolution's data-gradient epilogue."""
    return getattr(x, "_alignq_link", None) if (args.fuse_dgrad_bn and not args.sync_bn) else None


def conv_async_wgrad(x, weight, stride, padding, dilation, groups):
    if applies_stem(x, weight, stride, padding, dilation, groups, None):
        return stem_conv(x, weight)
    own = applies(x, weight, stride, padding, dilation, groups, None)
    own_w = own or applies_wgrad(x, weight, stride, padding, dilation, groups, None)
    own_d = own or applies_dgrad(x, weight, stride, padding, dilation, groups, None)
    return _ConvQFn.apply(x, weight, tuple(stride), tuple(padding), tuple(dilation), groups, own,
                          L.CONV_MODE_ID[args.own_conv] if (own_w or own_d) else 0, None, None, True, _up_link(x), own_w, own_d)


def conv_with_bn_stats(x, weight, bn, bn_ws):
    """Own 3x3 convolution whose epilogue also produces the batch statistics of ``bn`` (training mode): returns
    (conv output, save_mean, save_invstd).  The weight gradient goes to the side stream when ``args.async_wgrad``."""
    return _ConvQFn.apply(x, weight, (1, 1), (1, 1), (1, 1), 1, True, L.CONV_MODE_ID[args.own_conv], bn, bn_ws,
                          bool(args.async_wgrad), _up_link(x))
