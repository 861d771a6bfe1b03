"""``guassian_kernel`` / ``lmmd`` -- mirror of the reference's DSAN loss (cdf_alignment_admm/dsan_office/utils/mmd.py:9-41
with the label weights of utils/Weight.py:10-59) on the CUDA kernels of ``csrc/lmmd.cu``.

Same names (the reference's spelling ``guassian_kernel`` included), arguments and return values.  Differences: the label
weights are computed on the device with a handful of torch calls instead of numpy on the host (no ``.cpu()`` sync), and
the device is the inputs' (the reference hard-codes ``cuda:{args.gpus[0]}`` at import, mmd.py:7).
"""
from __future__ import annotations

import torch

from .. import _lib as L

__all__ = ["guassian_kernel", "lmmd", "cal_weight"]


def cal_weight(s_label, t_label, class_num=31):
    """``Weight.cal_weight(s_label, t_label, type='visual')`` (Weight.py:10-59) without leaving the device: returns
    (weight_ss, weight_tt, weight_st) as [B, B] fp32 tensors (all-zero when no class is shared, where the reference
    returns ``[0]``: the loss is 0 either way).  ``t_label`` are the target's class probabilities [B, class_num]."""
    B = s_label.shape[0]
    dev = t_label.device
    s_vec = torch.zeros(B, class_num, dtype=torch.float64, device=dev)
    s_vec[torch.arange(B, device=dev), s_label.to(dev).long()] = 1.0          # convert_to_onehot (Weight.py:4-5)
    s_sum = s_vec.sum(0, keepdim=True)
    in_s = s_sum.reshape(-1) > 0                                                # set_s
    s_sum = torch.where(s_sum == 0, torch.full_like(s_sum, 100.0), s_sum)
    s_vec = s_vec / s_sum
    t_vec = t_label.detach().to(torch.float32)
    t_sca = t_vec.max(1)[1]
    in_t = torch.zeros(class_num, dtype=torch.bool, device=dev)
    in_t[t_sca] = True                                                          # set_t: argmax labels of the target batch
    t_sum = t_vec.sum(0, keepdim=True)                                          # numpy float32 sum (Weight.py:22-23)
    t_sum = torch.where(t_sum == 0, torch.full_like(t_sum, 100.0), t_sum)
    t_vec = (t_vec / t_sum).double()                                            # np.dot promotes with the float64 source vec
    mask = (in_s & in_t).double()
    count = mask.sum()
    sm, tm = s_vec * mask, t_vec * mask
    denom = torch.where(count > 0, count, torch.ones_like(count))
    w_ss, w_tt, w_st = sm @ s_vec.t() / denom, tm @ t_vec.t() / denom, sm @ t_vec.t() / denom
    return w_ss.float(), w_tt.float(), w_st.float()


class _LmmdFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, total, W, kernel_mul, kernel_num, fix_sigma):
        tc = L.dev_f32(total, "lmmd features")
        Wc = L.dev_f32(W, "lmmd weights")
        n, d = tc.shape
        lib = L.load()
        ws = torch.zeros(int(lib.alignq_lmmd_ws_bytes(n)), dtype=torch.uint8, device=tc.device)
        coef = torch.empty(n, n, dtype=torch.float32, device=tc.device)
        loss = torch.empty(1, dtype=torch.float32, device=tc.device)
        with torch.cuda.device_of(tc):
            L.check(lib.alignq_lmmd_fwd(tc.data_ptr(), n, d, Wc.data_ptr(), float(kernel_mul), int(kernel_num),
                                        float(fix_sigma or 0.0), loss.data_ptr(), coef.data_ptr(), ws.data_ptr(), ws.numel(),
                                        L.stream_ptr()), "alignq_lmmd_fwd")
        ctx.save_for_backward(tc, coef, ws)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        tc, coef, ws = ctx.saved_tensors
        n, d = tc.shape
        g = torch.empty_like(tc)
        gl = L.dev_f32(gloss.reshape(1), "grad of lmmd loss")
        with torch.cuda.device_of(tc):
            L.check(L.load().alignq_lmmd_bwd(tc.data_ptr(), n, d, coef.data_ptr(), gl.data_ptr(), ws.data_ptr(), g.data_ptr(),
                                             L.stream_ptr()), "alignq_lmmd_bwd")
        return g, None, None, None, None


def guassian_kernel(source, target, kernel_mul=2.0, kernel_num=5, fix_sigma=None):
    """Sum of ``kernel_num`` Gaussian kernels of the pairwise squared distances of cat(source, target) (mmd.py:9-22),
    [2B, 2B].  Forward only as a stand-alone function (``lmmd`` is the differentiable entry)."""
    total = torch.cat([source, target], dim=0).detach()
    n = total.shape[0]
    diff = total.unsqueeze(0) - total.unsqueeze(1)          # [2B, 2B, d]: a diagnostic helper, not on the training path
    L2 = (diff * diff).sum(2)
    bw = fix_sigma if fix_sigma else torch.sum(L2) / (n * n - n)
    bw = bw / kernel_mul ** (kernel_num // 2)
    return sum(torch.exp(-L2 / (bw * kernel_mul ** i)) for i in range(kernel_num))


def lmmd(source, target, s_label, t_label, kernel_mul=2.0, kernel_num=5, fix_sigma=None, class_num=31):
    """``lmmd(source, target, s_label, t_label)`` (mmd.py:24-41): returns a 1-element tensor like the reference."""
    if not source.is_cuda:
        raise L.AlignQError("lmmd: expected CUDA tensors (alignq_b200 runs on sm_100a only; there is no CPU fallback)")
    B = source.shape[0]
    w_ss, w_tt, w_st = cal_weight(s_label, t_label, class_num)
    W = torch.cat([torch.cat([w_ss, -w_st], 1), torch.cat([-w_st.t(), w_tt], 1)], 0).contiguous()
    total = torch.cat([source, target], dim=0)
    return _LmmdFn.apply(total, W, kernel_mul, kernel_num, fix_sigma)
