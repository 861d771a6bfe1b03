"""GPU: fused BatchNorm2d -> activation quantizer -> (ReLU) against the un-fused
reference pipeline: torch BatchNorm2d -> oracle activation quantizer -> F.relu, forward, backward,
running statistics, eval mode.  The BN output differs from cuDNN's by an ulp (different summation order), so a
code that sits on a rounding tie may flip: bar = north_star's +-1 code on <= 1e-5 of the elements (at least 2
elements: one tie in a 100k-element tensor is already 1e-5), measured and printed per shape, beside the same count
for cuDNN's own fp32 BN against an fp64 BN (the reference's own tie floor).  Gradients: 1e-5 relative."""
import copy

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

import alignq_b200 as aq
from alignq_b200.model.fused import bn_act, can_fuse
from oracle import alignq_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def tie_budget(n):
    return max(2, int(1e-5 * n))


def close5(a, b, what):
    """1e-5 relative: |a - b| <= 1e-5 |b| + 1e-6 max|b|"""
    a, b = a.double(), b.double()
    d = (a - b).abs()
    tol = 1e-5 * b.abs() + 1e-6 * float(b.abs().max())
    assert bool((d <= tol).all()), f"{what}: worst |d|/tol {float((d / tol).max()):.2f}"


@pytest.mark.parametrize("variant", ["A", "B"])
@pytest.mark.parametrize("shape,relu", [((128, 16, 32, 32), True), ((64, 64, 8, 8), False), ((8, 24, 5, 7), True),
                                        ((4, 456, 8, 8), True), ((3, 1024, 2, 2), False)])
def test_fused_bn_act_matches_unfused_pipeline(variant, shape, relu):
    torch.manual_seed(0)
    aq.set_args(variant=variant, act_range=2, abitW=8, fuse_bn_act=True)
    B, C, H, W = shape
    x0 = (torch.randn(shape, device=DEV) * 1.5 + 0.3).contiguous(memory_format=torch.channels_last)
    gy = torch.randn(shape, device=DEV).contiguous(memory_format=torch.channels_last)
    bn = nn.BatchNorm2d(C).to(DEV).train()
    with torch.no_grad():
        bn.weight.copy_(1 + 0.2 * torch.randn(C))
        bn.bias.copy_(0.1 * torch.randn(C))
    bn_ref = copy.deepcopy(bn)
    actq = aq.activation_quantize_fn(8, "second")
    x = x0.clone().requires_grad_(True)
    assert can_fuse(bn, actq, x)
    y = bn_act(bn, actq, x, relu)
    (y * gy).sum().backward()
    xr = x0.clone().requires_grad_(True)
    z = bn_ref(xr)
    yr = O.activation_quantize(z, 8, "second", variant, 2.0)
    yr = F.relu(yr) if relu else yr
    (yr * gy).sum().backward()
    bn64 = copy.deepcopy(bn_ref).double()
    with torch.no_grad():                                          # bn_ref has stepped its running stats: irrelevant in train mode
        y64 = O.activation_quantize(bn64(x0.double()), 8, "second", variant, 2.0)
        y64 = F.relu(y64) if relu else y64
    n = 255
    step = (4.0 / n) if variant == "A" else (1.0 / n)
    d = (y - yr).abs()
    bad = int((d > 1e-6).sum())
    bad64 = int(((y.double() - y64).abs() > 1e-6).sum())
    floor64 = int(((yr.double() - y64).abs() > 1e-6).sum())
    print(f"bn_act {variant} {shape}: {bad}/{y.numel()} codes differ from cuDNN-BN + oracle ({bad / y.numel():.1e}); vs fp64 BN: "
          f"fused {bad64}, cuDNN fp32 {floor64}")
    assert float(d.max()) <= step * 1.001
    assert bad <= tie_budget(y.numel()), f"{bad} codes differ"
    assert bad64 <= max(tie_budget(y.numel()), 2 * floor64)
    assert y.is_contiguous(memory_format=torch.channels_last)
    same = d <= 1e-6                                               # a flipped code at 0 moves the ReLU mask of that element
    close5(x.grad * same, xr.grad * same, "gx")
    assert rel(x.grad, xr.grad) <= 1e-5 + 4.0 * bad / max(1.0, float(same.numel())) ** 0.5, "gx (all elements)"
    if bad == 0:
        close5(bn.weight.grad, bn_ref.weight.grad, "ggamma")
        close5(bn.bias.grad, bn_ref.bias.grad, "gbeta")
    else:                                                          # a flipped ReLU mask moves one addend of the sums
        assert rel(bn.weight.grad, bn_ref.weight.grad) <= 1e-4 and rel(bn.bias.grad, bn_ref.bias.grad) <= 1e-4
    assert torch.allclose(bn.running_mean, bn_ref.running_mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(bn.running_var, bn_ref.running_var, rtol=1e-5, atol=1e-6)
    assert int(bn.num_batches_tracked) == 1
    # eval mode: running statistics, no batch-statistics terms in the backward
    bn.eval(); bn_ref.eval()
    x2 = x0.clone().requires_grad_(True)
    y2 = bn_act(bn, actq, x2, relu)
    (y2 * gy).sum().backward()
    xr2 = x0.clone().requires_grad_(True)
    yr2 = O.activation_quantize(bn_ref(xr2), 8, "second", variant, 2.0)
    yr2 = F.relu(yr2) if relu else yr2
    (yr2 * gy).sum().backward()
    bad2 = int(((y2 - yr2).abs() > 1e-6).sum())
    assert bad2 <= tie_budget(y.numel())
    same2 = (y2 - yr2).abs() <= 1e-6
    close5(x2.grad * same2, xr2.grad * same2, "gx (eval mode)")


def test_fused_path_is_skipped_when_it_does_not_apply():
    aq.set_args(variant="A", act_range=2, fuse_bn_act=True)
    bn = nn.BatchNorm2d(16).to(DEV)
    actq = aq.activation_quantize_fn(8, "second")
    x_nchw = torch.randn(4, 16, 8, 8, device=DEV)
    assert not can_fuse(bn, actq, x_nchw)                                      # NCHW input -> separate modules
    assert not can_fuse(bn, aq.activation_quantize_fn(32, "second"), x_nchw.contiguous(memory_format=torch.channels_last))
    assert not can_fuse(bn, aq.activation_quantize_fn(8, "second", aq.ADMM(4)), x_nchw.contiguous(memory_format=torch.channels_last))
    aq.set_args(fuse_bn_act=False)
    assert not can_fuse(bn, actq, x_nchw.contiguous(memory_format=torch.channels_last))
    y = bn_act(bn, actq, x_nchw, True)                                         # still correct, un-fused
    ref = F.relu(O.activation_quantize(bn(x_nchw), 8, "second", "A", 2.0))
    assert int((y != ref).sum()) <= 1


def test_resnet20_fused_equals_unfused_within_band():
    from alignq_b200.model.resnet import resnet20_quant
    from oracle import models_oracle as MO
    torch.manual_seed(2)
    x = torch.randn(32, 3, 32, 32, device=DEV).contiguous(memory_format=torch.channels_last)
    outs = []
    for fuse in (False, True):
        aq.set_args(variant="A", bitW=8, abitW=8, act_range=2, fuse_bn_act=fuse)
        m = resnet20_quant(8, 8, "second")
        m.load_state_dict(MO.deterministic_fill(m.state_dict(), seed=5))
        m.to(DEV).train()
        out = m(x)
        out.logsumexp(1).sum().backward()
        gsum = torch.stack([p.grad.double().abs().sum() for p in m.parameters()])
        outs.append((out.detach(), gsum, m.bn.running_var.clone()))
    # same metrics and band as the model-level goldens: a BN output that differs in the last
    # ulp flips a few codes, and the 20-layer quantised net amplifies that like any 1-ulp perturbation
    assert rel(outs[1][0], outs[0][0]) <= 2e-2
    assert rel(outs[1][1], outs[0][1]) <= 5e-2
    assert torch.allclose(outs[1][2], outs[0][2], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("relu", [True, False])
def test_fused_bn_act_with_residual(relu):
    """Block tail `out = act_q1(bn1(conv1(out))); out += shortcut; out = F.relu(out)` (resnet.py:74-78) in one pass."""
    torch.manual_seed(3)
    aq.set_args(variant="A", act_range=2, abitW=8, fuse_bn_act=True)
    shape = (32, 32, 16, 16)
    cl = lambda t_: t_.contiguous(memory_format=torch.channels_last)
    x0, r0, gy = cl(torch.randn(shape, device=DEV)), cl(torch.randn(shape, device=DEV)), cl(torch.randn(shape, device=DEV))
    bn = nn.BatchNorm2d(shape[1]).to(DEV).train()
    bn_ref = copy.deepcopy(bn)
    actq = aq.activation_quantize_fn(8, "second")
    x, r = x0.clone().requires_grad_(True), r0.clone().requires_grad_(True)
    y = bn_act(bn, actq, x, relu, residual=r)
    (y * gy).sum().backward()
    xr, rr = x0.clone().requires_grad_(True), r0.clone().requires_grad_(True)
    yr = O.activation_quantize(bn_ref(xr), 8, "second", "A", 2.0) + rr
    yr = F.relu(yr) if relu else yr
    (yr * gy).sum().backward()
    d = (y - yr).abs()
    assert float(d.max()) <= 4.0 / 255 * 1.001 and int((d > 1e-6).sum()) <= tie_budget(y.numel())
    same = d <= 1e-6
    close5(x.grad * same, xr.grad * same, "gx")
    close5(r.grad * same, rr.grad * same, "g residual")
    assert rel(bn.weight.grad, bn_ref.weight.grad) <= 1e-4


@pytest.mark.parametrize("P,shape,relu,res", [(2, (64, 16, 16, 16), True, False), (4, (32, 64, 8, 8), True, True),
                                               (2, (16, 24, 5, 7), False, False)])
def test_sync_bn_entries_compose_to_the_single_device_kernels(P, shape, relu, res):
    """Data-parallel SyncBN inside the fused kernels (alignq_bn_act_sync_*): P ranks' row shards are run one after the
    other on ONE device, the all-reduce of the fp64 sums is a plain addition, and the result must equal the
    single-device fused kernels on the whole batch (same global statistics; +-1 code only at BN-output ties)."""
    from alignq_b200 import _lib as L
    from alignq_b200.model.fused import _bn_ws
    torch.manual_seed(7)
    aq.set_args(variant="A", act_range=2, abitW=8, fuse_bn_act=True, method="none")
    lib = L.load()
    cl = lambda t_: t_.contiguous(memory_format=torch.channels_last)
    B, C, H, W = shape
    x0, gy = cl(torch.randn(shape, device=DEV) * 1.4 + 0.2), cl(torch.randn(shape, device=DEV))
    r0 = cl(torch.randn(shape, device=DEV)) if res else None
    bn = nn.BatchNorm2d(C).to(DEV).train()
    with torch.no_grad():
        bn.weight.copy_(1 + 0.2 * torch.randn(C))
        bn.bias.copy_(0.1 * torch.randn(C))
    bn_s = copy.deepcopy(bn)
    actq = aq.activation_quantize_fn(8, "second")
    x = x0.clone().requires_grad_(True)
    rr = r0.clone().requires_grad_(True) if res else None
    y = bn_act(bn, actq, x, relu, residual=rr)
    (y * gy).sum().backward()
    # the same through the sync entries, shard by shard
    b = B // P
    rows, rows_g = b * H * W, B * H * W
    sh = lambda t_, r: cl(t_[r * b:(r + 1) * b])
    xs, gys = [sh(x0, r) for r in range(P)], [sh(gy, r) for r in range(P)]
    rs = [sh(r0, r) for r in range(P)] if res else [None] * P
    ws, counter = _bn_ws(bn_s, C, x0.device)
    st = L.stream_ptr()
    sums = []
    for r in range(P):
        s_ = torch.empty(2 * C, dtype=torch.float64, device=DEV)
        L.check(lib.alignq_bn_act_sync_stats(xs[r].data_ptr(), rows, C, s_.data_ptr(), ws.data_ptr(), counter.data_ptr(), st), "stats")
        sums.append(s_)
    tot = torch.stack(sums).sum(0)                                                        # the all-reduce
    ys, means, invs = [], [], []
    for r in range(P):
        yr = torch.empty_like(xs[r])
        m, iv = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
        rm, rv = bn_s.running_mean.clone(), bn_s.running_var.clone()                      # every rank updates its own copy
        L.check(lib.alignq_bn_act_sync_apply(xs[r].data_ptr(), rows, rows_g, C, tot.data_ptr(), bn_s.weight.data_ptr(),
                                             bn_s.bias.data_ptr(), rm.data_ptr(), rv.data_ptr(), float(bn_s.momentum),
                                             float(bn_s.eps), 8, 2.0, 0, int(relu), L.ptr(rs[r]), yr.data_ptr(), m.data_ptr(),
                                             iv.data_ptr(), 0, st), "apply")
        ys.append(yr); means.append(m); invs.append(iv)
    ysync = torch.cat(ys)
    assert int(((ysync - y.detach()).abs() > 1e-6).sum()) <= tie_budget(y.numel())
    assert torch.allclose(rm, bn.running_mean, rtol=1e-6, atol=1e-7) and torch.allclose(rv, bn.running_var, rtol=1e-6, atol=1e-7)
    bsums, gws, gbs = [], [], []
    for r in range(P):
        s_ = torch.empty(2 * C, dtype=torch.float64, device=DEV)
        gw, gb = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
        L.check(lib.alignq_bn_act_sync_bwd_reduce(xs[r].data_ptr(), ys[r].data_ptr(), gys[r].data_ptr(), rows, C,
                                                  bn_s.weight.data_ptr(), bn_s.bias.data_ptr(), means[r].data_ptr(),
                                                  invs[r].data_ptr(), 8, 2.0, 0, int(relu), s_.data_ptr(), gw.data_ptr(),
                                                  gb.data_ptr(), ws.data_ptr(), counter.data_ptr(), st), "bwd_reduce")
        bsums.append(s_); gws.append(gw); gbs.append(gb)
    btot = torch.stack(bsums).sum(0)
    gxs, grs = [], []
    for r in range(P):
        gx = torch.empty_like(xs[r])
        gr = torch.empty_like(xs[r]) if res else None
        L.check(lib.alignq_bn_act_sync_bwd_apply(xs[r].data_ptr(), ys[r].data_ptr(), gys[r].data_ptr(), rows, rows_g, C,
                                                 bn_s.weight.data_ptr(), bn_s.bias.data_ptr(), means[r].data_ptr(),
                                                 invs[r].data_ptr(), 8, 2.0, 0, int(relu), btot.data_ptr(), gx.data_ptr(),
                                                 L.ptr(gr), ws.data_ptr(), st), "bwd_apply")
        gxs.append(gx); grs.append(gr)
    same = (ysync - y.detach()).abs() <= 1e-6
    close5(torch.cat(gxs) * same, x.grad * same, "sync gx")
    assert rel(torch.stack(gws).sum(0), bn.weight.grad) <= 1e-5 and rel(torch.stack(gbs).sum(0), bn.bias.grad) <= 1e-5
    if res:
        close5(torch.cat(grs) * same, rr.grad * same, "sync g residual")


@pytest.mark.parametrize("shape", [(128, 16, 32, 32), (16, 32, 9, 7), (4, 64, 8, 8)])
def test_fork_sums_the_two_consumers_gradients_inside_the_backward_kernel(shape):
    """The input of a residual block has two consumers (resnet.py:70-78).  `fork` parks the second gradient and the
    producing bn-act backward reads gy + gy2 (alignq_bn_act_bwd_sum): same result as autograd's own accumulate kernel
    up to the rounding of (ga + gb) -- the kernel adds the same two fp32 values, so bit-equal -- and no add launch."""
    from alignq_b200.model.fused import fork
    torch.manual_seed(1)
    aq.set_args(variant="A", act_range=2, abitW=8, fuse_bn_act=True)
    B, C, H, W = shape
    x0 = (torch.randn(shape, device=DEV) * 1.2 - 0.1).contiguous(memory_format=torch.channels_last)
    wa = torch.randn(shape, device=DEV).contiguous(memory_format=torch.channels_last)
    wb = torch.randn(shape, device=DEV).contiguous(memory_format=torch.channels_last)
    bn = nn.BatchNorm2d(C).to(DEV).train()
    actq = aq.activation_quantize_fn(8, "second")
    grads = []
    for use_fork in (True, False):
        bn_i = copy.deepcopy(bn)
        x = x0.clone().requires_grad_(True)
        y = bn_act(bn_i, actq, x, True)
        assert hasattr(y, "_alignq_link")
        ya, yb = fork(y) if use_fork else (y, y)
        if use_fork:
            assert ya.data_ptr() == y.data_ptr() and yb.data_ptr() == y.data_ptr()
        ((ya * wa).sum() + (yb * yb * wb).sum()).backward()
        assert not y._alignq_link, "the parked gradient must have been consumed"
        grads.append((x.grad.clone(), bn_i.weight.grad.clone(), bn_i.bias.grad.clone()))
    for a, b, what in zip(grads[0], grads[1], ("gx", "ggamma", "gbeta")):
        assert torch.equal(a, b), f"{what}: max |d| {float((a - b).abs().max()):.3e}"
    with torch.no_grad():                                          # no graph: plain aliases
        y = bn_act(copy.deepcopy(bn), actq, x0, True)
        ya, yb = fork(y)
        assert ya is y and yb is y


@pytest.mark.parametrize("shape", [(128, 16, 32, 32), (8, 24, 5, 7)])
def test_single_launch_backward_equals_the_two_kernel_backward(shape, monkeypatch):
    """The cooperative one-launch backward (opt-in: ALIGNQ_BN_COOP_PER_SM, off by default since the weight-gradient
    side streams made it wait) must return what the default reduce + apply pair returns."""
    torch.manual_seed(2)
    aq.set_args(variant="A", act_range=2, abitW=8, fuse_bn_act=True)
    B, C, H, W = shape
    x0 = (torch.randn(shape, device=DEV) * 1.3 + 0.2).contiguous(memory_format=torch.channels_last)
    gy = torch.randn(shape, device=DEV).contiguous(memory_format=torch.channels_last)
    res = torch.randn(shape, device=DEV).contiguous(memory_format=torch.channels_last)
    bn = nn.BatchNorm2d(C).to(DEV).train()
    actq = aq.activation_quantize_fn(8, "second")
    outs = []
    for coop in ("0", "2"):
        monkeypatch.setenv("ALIGNQ_BN_COOP_PER_SM", coop)
        bn_i = copy.deepcopy(bn)
        x = x0.clone().requires_grad_(True)
        r = res.clone().requires_grad_(True)
        y = bn_act(bn_i, actq, x, True, residual=r)
        (y * gy).sum().backward()
        outs.append((x.grad.clone(), r.grad.clone(), bn_i.weight.grad.clone(), bn_i.bias.grad.clone()))
    for a, b, what in zip(outs[0], outs[1], ("gx", "g_residual", "ggamma", "gbeta")):
        close5(a, b, what)
