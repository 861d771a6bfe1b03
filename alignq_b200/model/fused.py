"""Fused BatchNorm2d -> activation quantizer -> (ReLU): the step either side of the quantizer in every
model file, e.g. ``F.relu(self.act_q0(self.bn0(out)))``
(cdf_alignment/resnet-20-cifar-10/model/resnet.py:72,121-123).

``bn_act(bn, actq, x, relu)`` keeps the reference's modules (same parameters, buffers and state_dict
keys) and only changes how they are executed: one statistics pass + one apply pass (12 B/elem)
instead of cuDNN BN + quantizer + ReLU (28 B/elem), and likewise backward.  It is used when
``args.fuse_bn_act`` is set and the input is a channels_last fp32 CUDA tensor with C % 4 == 0,
C <= 1024, a plain k-bit quantizer (no ADMM term); otherwise the three modules run one after the other.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib as L
from ..utils.options import args

def _bn_ws(bn, C: int, device):
    """(fp64 scratch, two zero-initialised uint32: ticket, epoch) of ONE BatchNorm layer.  The single-launch
    backward alternates between two accumulator sets and leaves the one it used dirty for the next launch on the
    same workspace to clear, so the scratch must not be shared between layers of different width."""
    ent = getattr(bn, "_alignq_bn_ws", None)
    need = int(L.load().alignq_bn_act_ws_doubles(C))
    if ent is None or ent[0].device != device or ent[0].numel() < need:
        ent = (torch.zeros(need, dtype=torch.float64, device=device), torch.zeros(2, dtype=torch.int32, device=device))
        bn._alignq_bn_ws = ent
    return ent


# ---- two consumers of one bn-act output without autograd's accumulate kernel ---------------------------------------------
# The input of a residual block feeds conv0 AND the shortcut (resnet.py:70-78), so autograd adds the two gradients with a
# stand-alone element-wise kernel before the producing bn-act backward runs: 9 launches of 2-7 us on the critical path of a
# ResNet-20 step.  `fork(x)` hands the block two aliases of x instead; in the backward pass the second alias' gradient
# is parked in the `link` the producing `_BnActFn` / `_PeerBnActFn` shares with its output, and that backward reads
# gy + gy2 straight from both buffers (alignq_bn_act_bwd_sum).
def _take_extra(link, like):
    """The parked second gradient of this layer's output (same layout as `like`), or None."""
    if link is None:
        return None
    extra = link.pop("extra", None)
    if extra is None:
        return None
    return L.like_layout(extra, like, "grad of the second consumer of a fused bn-act output")


class _GradFork(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, link):
        ctx.link = link
        ctx.set_materialize_grads(False)           # a parked shortcut gradient arrives as None, not as a zero tensor
        return x.detach(), x.detach()          # two aliases (not autograd views: a view output makes the engine re-materialise the gradient)

    @staticmethod
    def backward(ctx, ga, gb):
        if ga is None or gb is None:
            return (gb if ga is None else ga), None
        if (gb.shape != ga.shape or gb.stride() != ga.stride() or gb.dtype != torch.float32 or ga.dtype != torch.float32
                or "extra" in ctx.link):
            return ga + gb, None
        ctx.link["extra"] = gb                   # consumed by the producer's backward, which autograd runs next for ga
        return ga, None


def fork(x):
    """(x, x) for the two consumers of a block input -- first the convolution path, second the shortcut; aliases whose
    gradients are summed inside the producing bn-act backward when x came out of the fused bn-act path (see _GradFork),
    plain x twice otherwise."""
    link = getattr(x, "_alignq_link", None)
    if link is None or not (torch.is_grad_enabled() and x.requires_grad):
        return x, x
    xa, xb = _GradFork.apply(x, link)
    xa._alignq_link = link        # a convolution consuming xa may run the producer's backward reduce pass (conv_tc.py)
    xb._alignq_fork = link        # a bn-act layer taking xb as its residual parks that gradient in the link right away
    return xa, xb


class side_branch:
    """``with side_branch(x) as br: y = f(x)`` ... ``br.join(y)``: run an independent branch of the forward pass on a side
    stream beside what the caller launches until ``join`` (CUDA events both ways, so the pattern is captured into the
    step's CUDA graph as parallel branches; autograd replays each node's backward on its forward stream).  Active only
    inside a training step that opted in (``args.async_wgrad``) and never with cross-GPU BatchNorm statistics, whose
    kernels must all sit in one stream order; otherwise the branch simply runs in line."""

    _streams = {}

    def __init__(self, x):
        self.on = bool(args.async_wgrad and not args.sync_bn and x.is_cuda)
        self.ctx = None
        if self.on:
            dev = x.device.index
            if dev not in side_branch._streams:
                side_branch._streams[dev] = torch.cuda.Stream(device=x.device)
            self.side = side_branch._streams[dev]
            self.main = torch.cuda.current_stream(x.device)

    def __enter__(self):
        if self.on:
            self.side.wait_stream(self.main)
            self.ctx = torch.cuda.stream(self.side)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False

    def join(self, *tensors):
        if self.on:
            self.main.wait_stream(self.side)
            for t in tensors:
                t.record_stream(self.main)          # allocated on the side stream, consumed on the caller's


def _linked(fn, *fn_args, residual=None):
    """Run a bn-act autograd function with a fresh link and hang the link on its output (see fork)."""
    link = {}
    y = fn.apply(*fn_args, link, getattr(residual, "_alignq_fork", None))
    y._alignq_link = link
    return y


def _park_residual_grad(ctx, gr):
    """An identity shortcut taken from a fork: hand its gradient to the producer's link now (the convolution path's
    backward, which runs before the fork's, may want to add it in its epilogue) and return nothing to autograd."""
    if gr is not None and ctx.res_link is not None and "extra" not in ctx.res_link:
        ctx.res_link["extra"] = gr
        return None
    return gr


class _BnActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, bn, a_bit, act_range, variant, relu, residual=None, mean=None, invstd=None, link=None,
                res_link=None):
        B, C, H, W = x.shape
        rows = B * H * W
        training = bool(bn.training or bn.running_mean is None)
        y = torch.empty_like(x)
        have_stats = mean is not None          # the producing convolution's epilogue already reduced them
        if not have_stats:
            mean = torch.empty(C, dtype=torch.float32, device=x.device)
            invstd = torch.empty(C, dtype=torch.float32, device=x.device)
        ws, counter = _bn_ws(bn, C, x.device)
        ctx.bn = bn
        with torch.cuda.device_of(x):
            if have_stats:
                L.check(L.load().alignq_bn_act_apply(
                    x.data_ptr(), rows, C, L.ptr(weight), L.ptr(bias), mean.data_ptr(), invstd.data_ptr(), a_bit, act_range,
                    variant, int(relu), L.ptr(residual), y.data_ptr(), L.stream_ptr()), "alignq_bn_act_apply")
            else:
                L.check(L.load().alignq_bn_act_fwd(
                    x.data_ptr(), rows, C, L.ptr(weight), L.ptr(bias), L.ptr(bn.running_mean), L.ptr(bn.running_var),
                    float(bn.momentum), float(bn.eps), int(training), a_bit, act_range, variant, int(relu),
                    L.ptr(residual), y.data_ptr(), mean.data_ptr(), invstd.data_ptr(), ws.data_ptr(), counter.data_ptr(),
                    L.ptr(bn.num_batches_tracked) if training else 0, L.stream_ptr()), "alignq_bn_act_fwd")
        # y is only needed for the ReLU mask; without ReLU the model files go on to modify it in place
        # (`out += shortcut`, resnet.py:77), so it must not be saved
        ctx.save_for_backward(x, y if relu else None, weight, bias, mean, invstd)
        ctx.cfg = (rows, C, training, a_bit, act_range, variant, relu, residual is not None)
        ctx.link, ctx.res_link = link, res_link
        if link is not None and training:
            # what a data-gradient convolution needs to run this layer's backward reduce pass in its epilogue
            link["fwd"] = (x, mean, invstd, weight, bias, bn, (rows, C, a_bit, act_range, variant, relu))
        return y

    @staticmethod
    def backward(ctx, gy):
        x, y, weight, bias, mean, invstd = ctx.saved_tensors
        rows, C, training, a_bit, act_range, variant, relu, has_res = ctx.cfg
        gy = L.like_layout(gy, x, "grad of fused bn-act output")
        gy2 = _take_extra(ctx.link, x)
        red = ctx.link.pop("reduced", None) if ctx.link is not None else None
        if ctx.link is not None:
            ctx.link.pop("fwd", None)
        gx = torch.empty_like(x)
        gr = torch.empty_like(x) if (has_res and ctx.needs_input_grad[8]) else None
        ws, counter = _bn_ws(ctx.bn, C, x.device)
        if red is not None and gy2 is None and red[0] == gy.data_ptr():
            # the convolution that produced gy has run the reduce pass (and written the affine gradients): apply only
            with torch.cuda.device_of(x):
                L.check(L.load().alignq_bn_act_bwd_apply(
                    x.data_ptr(), L.ptr(y), gy.data_ptr(), rows, C, L.ptr(weight), L.ptr(bias), mean.data_ptr(),
                    invstd.data_ptr(), a_bit, act_range, variant, int(relu), gx.data_ptr(), L.ptr(gr), ws.data_ptr(),
                    L.stream_ptr()), "alignq_bn_act_bwd_apply")
            return gx, red[1], red[2], None, None, None, None, None, _park_residual_grad(ctx, gr), None, None, None, None
        gw = torch.empty(C, dtype=torch.float32, device=x.device) if weight is not None else None
        gb = torch.empty(C, dtype=torch.float32, device=x.device) if bias is not None else None
        with torch.cuda.device_of(x):
            L.check(L.load().alignq_bn_act_bwd_sum(
                x.data_ptr(), L.ptr(y), gy.data_ptr(), L.ptr(gy2), rows, C, L.ptr(weight), L.ptr(bias), mean.data_ptr(),
                invstd.data_ptr(), int(training), a_bit, act_range, variant, int(relu), gx.data_ptr(), L.ptr(gr),
                L.ptr(gw), L.ptr(gb), ws.data_ptr(), counter.data_ptr(), L.stream_ptr()), "alignq_bn_act_bwd_sum")
        return gx, gw, gb, None, None, None, None, None, _park_residual_grad(ctx, gr), None, None, None, None


class _SyncBnActFn(torch.autograd.Function):
    """Data-parallel form: BatchNorm statistics of the GLOBAL batch.  The kernels' own fp64 (sum, sum of squares)
    accumulators are all-reduced between the statistics and the apply launch (forward), and (sum g_z, sum g_z xhat)
    between the reduce and the apply launch (backward) -- one [2 C] fp64 NCCL all-reduce each, so ``sync_bn`` keeps the
    fused path instead of falling back to torch.nn.SyncBatchNorm + separate quantizer kernels."""

    @staticmethod
    def forward(ctx, x, weight, bias, bn, a_bit, act_range, variant, relu, residual, group, world):
        import torch.distributed as dist
        B, C, H, W = x.shape
        rows = B * H * W
        y = torch.empty_like(x)
        mean = torch.empty(C, dtype=torch.float32, device=x.device)
        invstd = torch.empty(C, dtype=torch.float32, device=x.device)
        ws, counter = _bn_ws(bn, C, x.device)
        sums = torch.empty(2 * C, dtype=torch.float64, device=x.device)
        lib = L.load()
        with torch.cuda.device_of(x):
            L.check(lib.alignq_bn_act_sync_stats(x.data_ptr(), rows, C, sums.data_ptr(), ws.data_ptr(), counter.data_ptr(),
                                                 L.stream_ptr()), "alignq_bn_act_sync_stats")
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
            rows_global = rows * world                             # equal shards (QATStep shards the batch evenly)
            L.check(lib.alignq_bn_act_sync_apply(
                x.data_ptr(), rows, rows_global, C, sums.data_ptr(), L.ptr(weight), L.ptr(bias), L.ptr(bn.running_mean),
                L.ptr(bn.running_var), float(bn.momentum), float(bn.eps), a_bit, act_range, variant, int(relu),
                L.ptr(residual), y.data_ptr(), mean.data_ptr(), invstd.data_ptr(), L.ptr(bn.num_batches_tracked),
                L.stream_ptr()), "alignq_bn_act_sync_apply")
        ctx.bn, ctx.group = bn, group
        ctx.save_for_backward(x, y if relu else None, weight, bias, mean, invstd)
        ctx.cfg = (rows, rows_global, C, a_bit, act_range, variant, relu, residual is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        import torch.distributed as dist
        x, y, weight, bias, mean, invstd = ctx.saved_tensors
        rows, rows_global, C, a_bit, act_range, variant, relu, has_res = ctx.cfg
        gy = L.like_layout(gy, x, "grad of fused bn-act output")
        gx = torch.empty_like(x)
        gr = torch.empty_like(x) if (has_res and ctx.needs_input_grad[8]) else None
        gw = torch.empty(C, dtype=torch.float32, device=x.device) if weight is not None else None
        gb = torch.empty(C, dtype=torch.float32, device=x.device) if bias is not None else None
        ws, counter = _bn_ws(ctx.bn, C, x.device)
        sums = torch.empty(2 * C, dtype=torch.float64, device=x.device)
        lib = L.load()
        with torch.cuda.device_of(x):
            L.check(lib.alignq_bn_act_sync_bwd_reduce(
                x.data_ptr(), L.ptr(y), gy.data_ptr(), rows, C, L.ptr(weight), L.ptr(bias), mean.data_ptr(), invstd.data_ptr(),
                a_bit, act_range, variant, int(relu), sums.data_ptr(), L.ptr(gw), L.ptr(gb), ws.data_ptr(), counter.data_ptr(),
                L.stream_ptr()), "alignq_bn_act_sync_bwd_reduce")
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=ctx.group)
            L.check(lib.alignq_bn_act_sync_bwd_apply(
                x.data_ptr(), L.ptr(y), gy.data_ptr(), rows, rows_global, C, L.ptr(weight), L.ptr(bias), mean.data_ptr(),
                invstd.data_ptr(), a_bit, act_range, variant, int(relu), sums.data_ptr(), gx.data_ptr(), L.ptr(gr),
                ws.data_ptr(), L.stream_ptr()), "alignq_bn_act_sync_bwd_apply")
        return gx, gw, gb, None, None, None, None, None, gr, None, None


class PeerExchange:
    """Peer-mapped buffers for the in-kernel SyncBN exchange (alignq_bn_act_*_peer): one symmetric-memory allocation
    per process group, rendezvoused once; the kernels write straight into the other ranks' copies over NVLink."""

    _inst = {}

    def __init__(self, group, device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.group = dist.group.WORLD if group is None else group
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        nbytes = int(L.load().alignq_bn_act_peer_bytes())
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.buf.zero_()
        self.seq = torch.zeros(2, dtype=torch.int32, device=device)   # exchange counter (+ one spare word)
        self.ptrs_dev = int(self.hdl.buffer_ptrs_dev)
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)                    # every rank's buffer is zero before anyone publishes

    @classmethod
    def get(cls, group, device):
        key = (id(group), device.index)
        if key not in cls._inst:
            cls._inst[key] = cls(group, device)
        return cls._inst[key]


class _PeerBnActFn(torch.autograd.Function):
    """Global-batch BatchNorm statistics with the exchange inside the kernels (NVLink peer stores + flags): still two
    launches forward and two backward, no NCCL call on the path."""

    @staticmethod
    def forward(ctx, x, weight, bias, bn, a_bit, act_range, variant, relu, residual, peer, link=None, res_link=None):
        B, C, H, W = x.shape
        rows = B * H * W
        rows_global = rows * peer.world
        y = torch.empty_like(x)
        mean = torch.empty(C, dtype=torch.float32, device=x.device)
        invstd = torch.empty(C, dtype=torch.float32, device=x.device)
        ws, counter = _bn_ws(bn, C, x.device)
        with torch.cuda.device_of(x):
            L.check(L.load().alignq_bn_act_fwd_peer(
                x.data_ptr(), rows, rows_global, C, L.ptr(weight), L.ptr(bias), L.ptr(bn.running_mean), L.ptr(bn.running_var),
                float(bn.momentum), float(bn.eps), a_bit, act_range, variant, int(relu), L.ptr(residual), y.data_ptr(),
                mean.data_ptr(), invstd.data_ptr(), ws.data_ptr(), counter.data_ptr(), L.ptr(bn.num_batches_tracked),
                peer.ptrs_dev, peer.seq.data_ptr(), peer.rank, peer.world, L.stream_ptr()), "alignq_bn_act_fwd_peer")
        ctx.bn, ctx.peer, ctx.link, ctx.res_link = bn, peer, link, res_link
        ctx.save_for_backward(x, y if relu else None, weight, bias, mean, invstd)
        ctx.cfg = (rows, rows_global, C, a_bit, act_range, variant, relu, residual is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, y, weight, bias, mean, invstd = ctx.saved_tensors
        rows, rows_global, C, a_bit, act_range, variant, relu, has_res = ctx.cfg
        peer = ctx.peer
        gy = L.like_layout(gy, x, "grad of fused bn-act output")
        gy2 = _take_extra(ctx.link, x)
        gx = torch.empty_like(x)
        gr = torch.empty_like(x) if (has_res and ctx.needs_input_grad[8]) else None
        gw = torch.empty(C, dtype=torch.float32, device=x.device) if weight is not None else None
        gb = torch.empty(C, dtype=torch.float32, device=x.device) if bias is not None else None
        ws, counter = _bn_ws(ctx.bn, C, x.device)
        with torch.cuda.device_of(x):
            L.check(L.load().alignq_bn_act_bwd_peer_sum(
                x.data_ptr(), L.ptr(y), gy.data_ptr(), L.ptr(gy2), rows, rows_global, C, L.ptr(weight), L.ptr(bias),
                mean.data_ptr(), invstd.data_ptr(), a_bit, act_range, variant, int(relu), gx.data_ptr(), L.ptr(gr), L.ptr(gw),
                L.ptr(gb), ws.data_ptr(), counter.data_ptr(), peer.ptrs_dev, peer.seq.data_ptr(), peer.rank, peer.world,
                L.stream_ptr()), "alignq_bn_act_bwd_peer_sum")
        return gx, gw, gb, None, None, None, None, None, _park_residual_grad(ctx, gr), None, None, None


def _sync_world():
    """(group, world) of the data-parallel SyncBN mode, or (None, 1)."""
    if not args.sync_bn:
        return None, 1
    from ..utils import dp_gram
    return dp_gram._state["group"], dp_gram.world()


def can_fuse(bn, actq, x) -> bool:
    return (bool(args.fuse_bn_act) and type(bn) is nn.BatchNorm2d and bn.momentum is not None
            and (bn.training or bn.running_mean is not None)
            and getattr(actq, "opt", None) is None and 1 <= actq.a_bit < 32
            and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4
            and x.shape[1] % 4 == 0 and x.shape[1] <= 1024
            and x.is_contiguous(memory_format=torch.channels_last) and x.data_ptr() % 16 == 0)


def conv_bn_act(conv, bn, actq, x, relu: bool, residual=None):
    """``relu(actq(bn(conv(x))) [+ residual])``: when the convolution runs on the own tcgen05 kernels and the bn-act pair is
    fused, the BatchNorm batch statistics come out of the convolution's epilogue (no statistics launch)."""
    if (args.own_conv != "off" and bn.training and conv.bias is None and not args.sync_bn
            and getattr(actq, "opt", None) is None and 1 <= actq.a_bit < 32 and type(bn) is nn.BatchNorm2d
            and bn.momentum is not None and bool(args.fuse_bn_act) and conv.out_channels in (16, 32)):
        from . import conv_tc
        weight_q = conv.quantize_fn(conv.weight)
        if residual is None and conv_tc.applies_stem(x, weight_q, conv.stride, conv.padding, conv.dilation, conv.groups, None):
            y_conv, mean, invstd = conv_tc.stem_conv(x, weight_q, bn, _bn_ws(bn, conv.out_channels, x.device))
            return _linked(_BnActFn, y_conv, bn.weight, bn.bias, bn, actq.a_bit, float(args.act_range),
                           L.VARIANT_ID[actq.variant], relu, None, mean, invstd)
        if (conv_tc.applies(x, weight_q, conv.stride, conv.padding, conv.dilation, conv.groups, None)
                and (residual is None or (residual.shape == x.shape and residual.stride() == x.stride()
                                          and residual.dtype == torch.float32 and residual.data_ptr() % 16 == 0))):
            y_conv, mean, invstd = conv_tc.conv_with_bn_stats(x, weight_q, bn, _bn_ws(bn, conv.out_channels, x.device))
            return _linked(_BnActFn, y_conv, bn.weight, bn.bias, bn, actq.a_bit, float(args.act_range),
                           L.VARIANT_ID[actq.variant], relu, residual, mean, invstd, residual=residual)
        # weight_q is already computed: finish the un-fused way without quantizing the weight twice
        if args.async_wgrad and x.is_cuda and conv.padding_mode == "zeros":
            y_conv = conv_tc.conv_async_wgrad(x, weight_q, conv.stride, conv.padding, conv.dilation, conv.groups)
        else:
            y_conv = F.conv2d(x, weight_q, None, conv.stride, conv.padding, conv.dilation, conv.groups)
        return bn_act(bn, actq, y_conv, relu, residual)
    return bn_act(bn, actq, conv(x), relu, residual)


def bn_act(bn, actq, x, relu: bool, residual=None):
    """``relu(actq(bn(x)) [+ residual])`` (or without the ReLU), fused when possible.  ``residual`` is the
    block shortcut of ``out += shortcut; out = F.relu(out)`` (resnet.py:77-78)."""
    if can_fuse(bn, actq, x) and (residual is None or (residual.shape == x.shape and residual.stride() == x.stride()
                                                       and residual.dtype == torch.float32
                                                       and residual.data_ptr() % 16 == 0)):
        group, world = _sync_world()
        if world > 1 and bn.training and args.sync_bn == "peer" and world <= 8:
            # global-batch statistics exchanged inside the kernels over NVLink peer memory
            return _linked(_PeerBnActFn, x, bn.weight, bn.bias, bn, actq.a_bit, float(args.act_range),
                           L.VARIANT_ID[actq.variant], relu, residual, PeerExchange.get(group, x.device), residual=residual)
        if world > 1 and bn.training:                      # global-batch statistics through an NCCL all-reduce
            return _SyncBnActFn.apply(x, bn.weight, bn.bias, bn, actq.a_bit, float(args.act_range),
                                      L.VARIANT_ID[actq.variant], relu, residual, group, world)
        return _linked(_BnActFn, x, bn.weight, bn.bias, bn, actq.a_bit, float(args.act_range),
                       L.VARIANT_ID[actq.variant], relu, residual, None, None, residual=residual)
    y = actq(bn(x))
    if residual is not None:
        y = y + residual
    return F.relu(y) if relu else y


# ---- classifier head + loss: avg_pool2d -> view -> linear -> cross_entropy in two launches (csrc/head_ce.cu) ---------------
_head_ws = {}


class _HeadCeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, weight, bias, target):
        B, C, H, W_ = feat.shape
        K = weight.shape[0]
        wc = L.dev_f32(weight, "classifier weight")
        bc = L.dev_f32(bias, "classifier bias") if bias is not None else None
        dev = feat.device
        key = (dev.index, B)                       # ticket + per-sample losses: one head at a time per device
        ws = _head_ws.get(key)
        if ws is None:
            ws = torch.zeros(int(L.load().alignq_head_ce_ws_bytes(B)), dtype=torch.uint8, device=dev)
            _head_ws[key] = ws
        logits = torch.empty(B, K, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        pooled = torch.empty(B, C, dtype=torch.float32, device=dev)
        glog = torch.empty(B, K, dtype=torch.float32, device=dev)
        with torch.cuda.device_of(feat):
            L.check(L.load().alignq_head_ce_fwd(feat.data_ptr(), wc.data_ptr(), L.ptr(bc), target.data_ptr(), B, H * W_, C, K,
                                                logits.data_ptr(), loss.data_ptr(), pooled.data_ptr(), glog.data_ptr(),
                                                ws.data_ptr(), ws.numel(), L.stream_ptr()), "alignq_head_ce_fwd")
        ctx.save_for_backward(wc, pooled, glog)
        ctx.geom = (B, H, W_, C, K)
        ctx.like = feat
        ctx.has_bias = bias is not None
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(logits)
        return loss, logits

    @staticmethod
    def backward(ctx, gloss, _glogits=None):
        if gloss is None:
            return None, None, None, None
        wc, pooled, glog = ctx.saved_tensors
        B, H, W_, C, K = ctx.geom
        gl = L.dev_f32(gloss.reshape(1), "grad of the loss")
        gf = torch.empty_like(ctx.like) if ctx.needs_input_grad[0] else None       # channels_last like feat
        gw = torch.empty_like(wc) if ctx.needs_input_grad[1] else None
        gb = torch.empty(K, dtype=torch.float32, device=wc.device) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        with torch.cuda.device_of(wc):
            L.check(L.load().alignq_head_ce_bwd(wc.data_ptr(), pooled.data_ptr(), glog.data_ptr(), gl.data_ptr(), B, H * W_, C, K,
                                                L.ptr(gf), L.ptr(gw), L.ptr(gb), L.stream_ptr()), "alignq_head_ce_bwd")
        return gf, gw, gb, None


def head_ce_applies(feat, linear, target) -> bool:
    return (feat.is_cuda and feat.dtype == torch.float32 and feat.dim() == 4 and type(linear) is nn.Linear
            and feat.shape[1] % 4 == 0 and feat.shape[1] <= 1024 and linear.out_features <= 1024
            and linear.in_features == feat.shape[1] and feat.data_ptr() % 16 == 0
            and (feat.is_contiguous(memory_format=torch.channels_last) or feat.shape[2] * feat.shape[3] == 1)
            and target.dtype == torch.int64 and target.dim() == 1 and target.is_contiguous())


def avgpool_linear_ce(feat, linear, target):
    """``F.cross_entropy(linear(F.avg_pool2d(feat, feat.size(3)).view(B, -1)), target)`` (model/resnet.py:108-110 +
    main.py:283-286) on the two kernels of csrc/head_ce.cu: returns (loss, logits); logits carry no gradient (they are
    what the reference's accuracy meters read)."""
    if head_ce_applies(feat, linear, target):
        return _HeadCeFn.apply(feat, linear.weight, linear.bias, target)
    logits = linear(F.adaptive_avg_pool2d(feat, 1).view(feat.size(0), -1))
    return F.cross_entropy(logits, target), logits.detach()
