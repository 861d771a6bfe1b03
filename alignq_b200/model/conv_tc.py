"""``F.conv2d(input, weight_q, None, 1, 1)`` of ``Conv2d_Q.forward`` (QA:116-120) on the hand-written tcgen05 kernels
(csrc/conv_tc.cu): forward, data gradient and weight gradient of the 3x3 / stride 1 / padding 1 / Cin == Cout in
{16, 32, 64} convolutions on channels_last fp32 tensors -- SURVEY.md 8(f) item 2.  Every other convolution of the model
files (stem, strided, 1x1 projection, depthwise, NCHW inputs) keeps calling the library convolution, as the reference does.

``args.own_conv``: "off" (library convolution everywhere), "tf32" (one tensor-core pass on tf32 operands: the numerics of
cuDNN under torch's default ``allow_tf32 = True``) or "tf32x3" (three passes on H + L split operands: fp32 parity, 1e-5).
"""
from __future__ import annotations

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


def applies(x, weight, stride, padding, dilation, groups, bias) -> bool:
    if args.own_conv == "off" or bias is not None or groups != 1:
        return False
    if tuple(stride) != (1, 1) or tuple(padding) != (1, 1) or tuple(dilation) != (1, 1):
        return False
    if weight.dim() != 4 or tuple(weight.shape[2:]) != (3, 3) or weight.shape[0] != weight.shape[1]:
        return False
    C = weight.shape[0]
    if C not in (16, 32, 64) or C not in tuple(args.own_conv_channels) or (C == 64 and args.own_conv == "tf32x3"):
        return False
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == C
            and x.is_contiguous(memory_format=torch.channels_last) and x.data_ptr() % 16 == 0
            and weight.dtype == torch.float32 and weight.data_ptr() % 16 == 0)


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
