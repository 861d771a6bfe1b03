"""GPU: the tcgen05 3x3 convolution kernels (csrc/conv_tc.cu) against F.conv2d in fp32 (TF32 off):
forward, data gradient and weight gradient.  Bars: TF32X3 mode (three passes on H + L split operands) 1e-5 relative
to the output's magnitude -- fp32 parity; TF32 mode (one pass, the numerics cuDNN uses under torch's default
allow_tf32 = True) 2e-3, with cuDNN's own TF32 error on the same inputs printed beside it."""
import pytest
import torch
import torch.nn.functional as F

import alignq_b200 as aq
from alignq_b200 import _lib as L
from alignq_b200.model import conv_tc

pytestmark = pytest.mark.gpu
DEV = "cuda"
SHAPES = [(2, 8, 8, 16), (128, 32, 32, 16), (128, 16, 16, 32), (128, 8, 8, 64), (3, 5, 7, 32), (1, 32, 32, 16), (5, 3, 3, 16),
          (7, 14, 14, 64), (130, 32, 32, 16)]


@pytest.fixture(autouse=True)
def _fp32_reference():
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old
    aq.reset_args()


def relmax(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


@pytest.mark.parametrize("mode,tol", [("tf32x3", 1e-5), ("tf32", 2e-3)])
@pytest.mark.parametrize("N,H,W,C", SHAPES)
def test_conv3x3_forward_backward_vs_fp32_conv2d(mode, tol, N, H, W, C):
    if C == 64 and mode == "tf32x3":
        pytest.skip("C = 64 in TF32X3 mode is not covered (split weights exceed shared memory): library convolution")
    torch.manual_seed(50)
    aq.set_args(own_conv=mode, own_conv_channels=(16, 32, 64))
    x0 = torch.randn(N, C, H, W, device=DEV).contiguous(memory_format=torch.channels_last)
    w0 = (torch.randn(C, C, 3, 3, device=DEV) * (2.0 / (9 * C)) ** 0.5).contiguous(memory_format=torch.channels_last)
    gy = torch.randn(N, C, H, W, device=DEV).contiguous(memory_format=torch.channels_last)
    assert conv_tc.applies(x0, w0, (1, 1), (1, 1), (1, 1), 1, None)
    x, w = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
    y = conv_tc.conv3x3(x, w)
    (y * gy).sum().backward()
    xr, wr = x0.double().clone().requires_grad_(True), w0.double().clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, None, 1, 1)
    (yr * gy.double()).sum().backward()
    x32, w32 = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
    y32 = F.conv2d(x32, w32, None, 1, 1)
    (y32 * gy).sum().backward()
    e = (relmax(y, yr), relmax(x.grad, xr.grad), relmax(w.grad, wr.grad))
    f = (relmax(y32, yr), relmax(x32.grad, xr.grad), relmax(w32.grad, wr.grad))
    print(f"conv3x3 {mode} N={N} {H}x{W} C={C}: max|d|/max|ref| fwd {e[0]:.1e} dgrad {e[1]:.1e} wgrad {e[2]:.1e}   "
          f"(cuDNN fp32: {f[0]:.1e} {f[1]:.1e} {f[2]:.1e})")
    assert y.is_contiguous(memory_format=torch.channels_last) and x.grad.shape == x0.shape and w.grad.stride() == w0.stride()
    assert e[0] <= tol and e[1] <= tol and e[2] <= tol


def test_conv2d_q_uses_the_kernels_when_it_applies_and_the_library_otherwise():
    torch.manual_seed(51)
    aq.set_args(variant="A", bitW=8, own_conv="tf32x3", own_conv_channels=(16, 32))
    lib = L.load()
    conv = aq.conv2d_Q_fn(8, "second")(32, 32, 3, padding=1, bias=False).to(DEV)
    conv.weight.data = conv.weight.data.contiguous(memory_format=torch.channels_last)
    x = torch.randn(8, 32, 16, 16, device=DEV).contiguous(memory_format=torch.channels_last)
    n0 = lib.alignq_launch_count()
    y = conv(x)
    assert lib.alignq_launch_count() - n0 >= 3                        # weight quantizer (2) + own conv
    aq.set_args(own_conv="off")
    y_lib = conv(x)
    assert relmax(y, y_lib) <= 1e-5
    aq.set_args(own_conv="tf32x3")
    for bad in (torch.randn(8, 32, 16, 16, device=DEV),                # NCHW input
                ):
        n1 = lib.alignq_launch_count()
        conv(bad)
        assert lib.alignq_launch_count() - n1 == 2                    # only the weight quantizer: library convolution
    strided = aq.conv2d_Q_fn(8, "second")(32, 32, 3, stride=2, padding=1, bias=False).to(DEV)
    assert not conv_tc.applies(x, strided.weight, strided.stride, strided.padding, strided.dilation, 1, None)


@pytest.mark.parametrize("C,H", [(16, 32), (32, 16)])
def test_conv_epilogue_bn_statistics_match_the_bn_act_statistics_pass(C, H):
    """alignq_conv3x3_fwd_bnstats + alignq_bn_act_apply == alignq_conv3x3_fwd + alignq_bn_act_fwd: the BatchNorm batch
    statistics taken from the convolution's epilogue (one launch fewer per layer) against the fused kernels' own statistics
    pass -- save_mean / save_invstd / running statistics to fp32 round-off, activation codes equal up to BN-output ties."""
    import copy
    from alignq_b200.model.fused import bn_act, conv_bn_act
    torch.manual_seed(52)
    aq.set_args(variant="A", bitW=8, abitW=8, act_range=2, fuse_bn_act=True, own_conv="tf32x3", own_conv_channels=(16, 32),
                method="none")
    conv = aq.conv2d_Q_fn(8, "second")(C, C, 3, padding=1, bias=False).to(DEV)
    conv.weight.data = conv.weight.data.contiguous(memory_format=torch.channels_last)
    bn = torch.nn.BatchNorm2d(C).to(DEV).train()
    with torch.no_grad():
        bn.weight.copy_(1 + 0.2 * torch.randn(C))
        bn.bias.copy_(0.1 * torch.randn(C))
    bn2, conv2 = copy.deepcopy(bn), copy.deepcopy(conv)
    q = aq.activation_quantize_fn(8, "second")
    x0 = torch.randn(64, C, H, H, device=DEV).contiguous(memory_format=torch.channels_last)
    r0 = torch.randn(64, C, H, H, device=DEV).contiguous(memory_format=torch.channels_last)
    gy = torch.randn_like(x0)
    x, r = x0.clone().requires_grad_(True), r0.clone().requires_grad_(True)
    y = conv_bn_act(conv, bn, q, x, True, residual=r)                                # statistics from the conv epilogue
    (y * gy).sum().backward()
    x2, r2 = x0.clone().requires_grad_(True), r0.clone().requires_grad_(True)
    y2 = bn_act(bn2, q, conv2(x2), True, residual=r2)                                # conv, then stats + apply launches
    (y2 * gy).sum().backward()
    assert torch.allclose(bn.running_mean, bn2.running_mean, rtol=1e-5, atol=1e-7)
    assert torch.allclose(bn.running_var, bn2.running_var, rtol=1e-5, atol=1e-7)
    assert int(bn.num_batches_tracked) == 1
    bad = int(((y - y2).abs() > 1e-6).sum())
    assert bad <= max(2, int(1e-5 * y.numel())), f"{bad} codes differ"
    same = (y - y2).abs() <= 1e-6
    assert relmax(x.grad, x2.grad) <= 1e-4 and relmax(conv.weight.grad, conv2.weight.grad) <= 1e-4
    assert relmax(r.grad * same, r2.grad * same) <= 1e-6
    assert relmax(bn.weight.grad, bn2.weight.grad) <= 1e-4


# ---- first-layer (Cin = 3) convolution: csrc/conv_stem.cu --------------------------------------------------------------
@pytest.mark.parametrize("N,H,W,Cout", [(128, 32, 32, 16), (3, 32, 32, 32), (5, 28, 28, 16), (2, 17, 23, 16), (1, 3, 3, 32),
                                         (7, 40, 70, 16), (300, 8, 8, 16)])
def test_stem_conv_forward_and_weight_gradient_vs_fp64_conv2d(N, H, W, Cout):
    """conv0 = Conv2d(3, Cout, 3, 1, 1) (model/resnet.py:92 through quantization.py:116-120): direct fp32 kernels, exact
    fp32 products -- 2e-6 of max|ref| against the fp64 convolution (cuDNN's fp32 path printed beside it); repeated calls
    exercise the re-armed fp64 accumulators of the weight gradient."""
    torch.manual_seed(60)
    aq.set_args(own_conv="tf32")
    x0 = torch.randn(N, 3, H, W, device=DEV).contiguous(memory_format=torch.channels_last)
    w0 = (torch.randn(Cout, 3, 3, 3, device=DEV) * (2.0 / 27) ** 0.5).contiguous(memory_format=torch.channels_last)
    gy = torch.randn(N, Cout, H, W, device=DEV).contiguous(memory_format=torch.channels_last)
    assert conv_tc.applies_stem(x0, w0, (1, 1), (1, 1), (1, 1), 1, None)
    xr, wr = x0.double(), w0.double().clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, None, 1, 1)
    (yr * gy.double()).sum().backward()
    w32 = w0.clone().requires_grad_(True)
    y32 = F.conv2d(x0, w32, None, 1, 1)
    (y32 * gy).sum().backward()
    for rep in range(3):
        w = w0.clone().requires_grad_(True)
        y = conv_tc.stem_conv(x0, w)
        (y * gy).sum().backward()
        e = (relmax(y, yr), relmax(w.grad, wr.grad))
        assert y.shape == (N, Cout, H, W) and y.is_contiguous(memory_format=torch.channels_last)
        assert w.grad.shape == w0.shape and w.grad.stride() == w0.stride()
        assert e[0] <= 2e-6 and e[1] <= 2e-6, e
    print(f"stem conv N={N} {H}x{W} Cout={Cout}: max|d|/max|ref| fwd {e[0]:.1e} wgrad {e[1]:.1e}   "
          f"(cuDNN fp32: {relmax(y32, yr):.1e} {relmax(w32.grad, wr.grad):.1e})")


def test_stem_conv_declines_what_it_does_not_cover():
    aq.set_args(own_conv="tf32")
    x = torch.randn(2, 3, 8, 8, device=DEV).contiguous(memory_format=torch.channels_last)
    w = torch.randn(16, 3, 3, 3, device=DEV)
    assert conv_tc.applies_stem(x, w, (1, 1), (1, 1), (1, 1), 1, None)
    assert not conv_tc.applies_stem(x, w, (2, 2), (1, 1), (1, 1), 1, None)                       # strided
    assert not conv_tc.applies_stem(x.clone().requires_grad_(True), w, (1, 1), (1, 1), (1, 1), 1, None)   # no data gradient
    assert not conv_tc.applies_stem(x.contiguous(), w, (1, 1), (1, 1), (1, 1), 1, None)          # NCHW
    assert not conv_tc.applies_stem(x, torch.randn(24, 3, 3, 3, device=DEV), (1, 1), (1, 1), (1, 1), 1, None)
    aq.set_args(own_conv="off")
    assert not conv_tc.applies_stem(x, w, (1, 1), (1, 1), (1, 1), 1, None)
    lib = L.load()
    assert lib.alignq_conv3x3_stem_fwd(x.data_ptr(), w.data_ptr(), x.data_ptr(), 2, 8, 8, 24, 0, 0, 0.0, 0.0, 0, 0, 0, 0, 0, 0) == -3      # ALIGNQ_ERANGE


@pytest.mark.parametrize("Cout,H", [(16, 32), (32, 12)])
def test_stem_conv_epilogue_bn_statistics(Cout, H):
    """conv_bn_act on conv0: BatchNorm batch statistics from the stem kernel's epilogue vs the statistics launch."""
    import copy
    from alignq_b200.model.fused import bn_act, conv_bn_act
    torch.manual_seed(61)
    aq.set_args(variant="A", bitW=8, abitW=8, act_range=2, fuse_bn_act=True, own_conv="tf32", method="none")
    conv = aq.conv2d_Q_fn(8, "second")(3, Cout, 3, padding=1, bias=False).to(DEV)
    conv.weight.data = conv.weight.data.contiguous(memory_format=torch.channels_last)
    bn = torch.nn.BatchNorm2d(Cout).to(DEV).train()
    with torch.no_grad():
        bn.weight.copy_(1 + 0.2 * torch.randn(Cout))
        bn.bias.copy_(0.1 * torch.randn(Cout))
    bn2, conv2 = copy.deepcopy(bn), copy.deepcopy(conv)
    q = aq.activation_quantize_fn(8, "second")
    x = torch.randn(64, 3, H, H, device=DEV).contiguous(memory_format=torch.channels_last)
    gy = torch.randn(64, Cout, H, H, device=DEV).contiguous(memory_format=torch.channels_last)
    for rep in range(2):
        y = conv_bn_act(conv, bn, q, x, True)
        (y * gy).sum().backward()
    aq.set_args(own_conv="off")
    for rep in range(2):
        y2 = bn_act(bn2, q, conv2(x), True)
        (y2 * gy).sum().backward()
    assert int(bn.num_batches_tracked) == 2
    assert torch.allclose(bn.running_mean, bn2.running_mean, rtol=1e-4, atol=1e-6)
    assert torch.allclose(bn.running_var, bn2.running_var, rtol=1e-4, atol=1e-6)
    bad = int(((y - y2).abs() > 1e-6).sum())
    assert bad <= max(2, int(1e-4 * y.numel())), f"{bad} codes differ"      # BN-output rounding ties only (fp32 library arm)
    assert relmax(conv.weight.grad, conv2.weight.grad) <= 1e-3
    assert relmax(bn.weight.grad, bn2.weight.grad) <= 1e-3


@pytest.mark.parametrize("C,H", [(16, 32)])
@pytest.mark.parametrize("mode", ["tf32x3", "tf32"])
def test_bn_act_backward_reduce_in_the_data_gradient_epilogue(mode, C, H):
    """Two residual blocks of the C = 16 stage (resnet.py:51-55): with ``fuse_dgrad_bn`` the data-gradient kernel of each
    own convolution also runs the reduce pass of the preceding bn-act layer's backward (sum g_z, sum g_z xhat, the
    affine gradients) and adds the parked shortcut gradient, so that layer's backward is its apply pass alone
    (alignq_conv3x3_bwd_data_bnreduce + alignq_bn_act_bwd_apply).  Same forward, so the gradients must agree with the
    un-fused backward to fp32 summation-order noise."""
    import copy
    from alignq_b200.model.fused import bn_act
    from alignq_b200.model.resnet import PreActBlock_conv_Q
    torch.manual_seed(70)
    base = dict(variant="A", bitW=8, abitW=8, act_range=2, fuse_bn_act=True, own_conv=mode, own_conv_channels=(16, 32), method="none")
    aq.set_args(**base)
    N = 32
    blocks = torch.nn.ModuleList([PreActBlock_conv_Q("second", 8, 8, C, C, 1, variant="A") for _ in range(2)]).to(DEV).train()
    bn0 = torch.nn.BatchNorm2d(C).to(DEV).train()
    q0 = aq.activation_quantize_fn(8, "second")
    for m in blocks.modules():
        if isinstance(m, torch.nn.Conv2d):
            m.weight.data = m.weight.data.contiguous(memory_format=torch.channels_last)
        if isinstance(m, torch.nn.BatchNorm2d):
            with torch.no_grad():
                m.weight.copy_(1 + 0.2 * torch.randn(C, device=DEV))
                m.bias.copy_(0.1 * torch.randn(C, device=DEV))
    x0 = torch.randn(N, C, H, H, device=DEV).contiguous(memory_format=torch.channels_last)
    gy = torch.randn_like(x0)
    lib = L.load()
    res = []
    for fuse in (True, False):
        aq.set_args(fuse_dgrad_bn=fuse)
        net, bn = copy.deepcopy(blocks), copy.deepcopy(bn0)
        x = x0.clone().requires_grad_(True)
        n0 = lib.alignq_launch_count()
        out = bn_act(bn, q0, x, True)               # the stem's bn-act: its output feeds block 1 (conv path + shortcut)
        for blk in net:
            out = blk(out)
        (out * gy).sum().backward()
        torch.cuda.synchronize()
        grads = [x.grad.clone()] + [p.grad.clone() for p in list(bn.parameters()) + list(net.parameters())]
        res.append((out.detach().clone(), grads, lib.alignq_launch_count() - n0))
    aq.set_args(fuse_dgrad_bn=True)
    assert torch.equal(res[0][0], res[1][0])                          # identical forward
    assert res[0][2] == res[1][2] - 4, (res[0][2], res[1][2])         # four reduce launches fewer (one per own convolution)
    worst = 0.0
    for a, b in zip(res[0][1], res[1][1]):
        assert a.shape == b.shape
        worst = max(worst, relmax(a, b))
    print(f"fuse_dgrad_bn {mode}: worst max|d|/max|ref| over x.grad and {len(res[0][1]) - 1} parameter gradients {worst:.2e}")
    assert worst <= 2e-5, worst
