"""GPU (one device): the CUDA backend of the data-parallel global-batch ADMM term (alignq_b200/utils/dp_gram.py).
The collectives are exercised over gloo on CPU (tests/test_dp_gram_gloo.py) and over NCCL on 2+ GPUs by
tools/dp_parity_check.py / `bench.py --gpus N --dp-parity`; here the per-rank kernels are composed by hand for
P = 2 and 4 feature slices on ONE device and compared with the single-device fused path and the fp64 oracle:
  sum over slices of alignq_gram_sums_fwd  -> alignq_gram_sums_to_d  ==  D of alignq_act_admm_fwd
  concat over slices of alignq_act_admm_bwd(gy = NULL, gloss / P), then alignq_act_bwd_add  ==  the fused backward."""
import pytest
import torch

import alignq_b200 as aq
from alignq_b200 import _lib as L
from alignq_b200.utils.dp_gram import CudaBackend
from oracle import alignq_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("mode", ["fp32", "tf32x3"])
@pytest.mark.parametrize("P,B,shape,eps", [(2, 128, (16, 16, 16), 0.0), (4, 28, (64, 14, 14), 1e-5), (2, 100, (8, 6, 6), 0.0),
                                           (8, 128, (64, 8, 8), 0.0), (2, 160, (4, 8, 8), 0.0)])
def test_feature_slices_compose_to_the_single_device_term(mode, P, B, shape, eps):
    torch.manual_seed(40)
    variant = "B" if eps == 0.0 else "C"
    aq.set_args(variant=variant, act_range=2, method="ours", gram_mode=mode, dp_gram="replica")
    be = CudaBackend()
    gm = L.GRAM_MODE_ID[mode]
    x0 = torch.randn(B, *shape, device=DEV)
    gy = torch.randn_like(x0)
    F = x0[0].numel()
    Fs = F // P
    dim = B
    admm = aq.ADMM(dim).to(DEV)
    Z, U = admm.alterD.detach(), admm.gamma.detach()
    # single-device fused path
    Fn = aq.activation_quantize_fn if variant == "B" else aq.activation_quantize_fn2
    xr = x0.clone().requires_grad_(True)
    y_ref, loss_ref = Fn(8, "second", admm)(xr)
    D_ref = admm.D.clone()
    w = 1.3
    ((y_ref * gy).sum() + w * loss_ref).backward()
    # hand-composed feature-sharded path
    x2d = x0.view(B, F)
    slices = [x2d[:, r * Fs:(r + 1) * Fs].contiguous() for r in range(P)]
    sums = sum(be.gram_sums(s, 2.0, eps, gm) for s in slices)                     # the all-reduce
    D, loss, dLdD = be.admm_from_sums(sums, F, Z, U, 0.2, 0.3)
    gmax = float(O.corr(x2d.double(), x2d.double(), eps).abs().max())
    assert float((D - D_ref).abs().max()) <= 2e-6 * gmax, "D from feature slices"
    assert abs(float(loss) - float(loss_ref)) <= 1e-5 * abs(float(loss_ref))
    gl = torch.full((1,), w, device=DEV)
    g_slices = [be.slice_bwd(s, dLdD, gl / P, 8, 2.0, eps, gm) for s in slices]
    gadd = torch.cat(g_slices, dim=1).view_as(x0)                                  # the all-to-all back
    gx = be.act_bwd_add(x0, gy, gadd, 8, 2.0, L.VARIANT_ID[variant])
    assert torch.equal(be.act_fwd(x0, 8, 2.0, L.VARIANT_ID[variant]), y_ref.detach())
    # against fp64 autograd of the oracle on the whole batch: north_star's 1e-5
    x64 = x0.double().clone().requires_grad_(True)
    y64, l64, D64 = O.activation_quantize_admm(x64, 8, Z.double(), U.double(), "second", variant, 2.0)
    ((y64 * gy.double()).sum() + w * l64).backward()
    d = (gx.double() - x64.grad).abs()
    tol = 1e-5 * x64.grad.abs() + 1e-6 * float(x64.grad.abs().max())
    assert bool((d <= tol).all()), f"gx: worst |d|/tol {float((d / tol).max()):.2f}"
    d2 = (gx - xr.grad).abs()
    assert float(d2.max()) <= 1e-5 * float(xr.grad.abs().max())
    # the pure ADMM part alone (no gy) is what travels through the all-to-all: 1e-5 of its own magnitude
    x64b = x0.double().clone().requires_grad_(True)
    _, l64b, _ = O.activation_quantize_admm(x64b, 8, Z.double(), U.double(), "second", variant, 2.0)
    (w * l64b).backward()
    e = float((gadd.double() - x64b.grad).abs().max() / x64b.grad.abs().max())
    print(f"dp slices P={P} B={B} {mode}: pure ADMM gradient max|d|/max|ref| = {e:.2e}")
    assert e <= 1e-5


def test_act_bwd_add_matches_separate_kernels():
    torch.manual_seed(41)
    be = CudaBackend()
    for n in (1, 5, 4099, 1 << 20):
        x, gy, ga = (torch.randn(n, device=DEV) for _ in range(3))
        out = be.act_bwd_add(x, gy, ga, 8, 2.0, 1)
        ref = ga + be.act_bwd(x, gy, 8, 2.0, 1)
        assert torch.equal(out, ref)
    base = torch.randn(4100, device=DEV)
    x, gy, ga = base[1:], torch.randn(4099, device=DEV), torch.randn(4099, device=DEV)     # unaligned view
    assert torch.equal(be.act_bwd_add(x, gy, ga, 8, 2.0, 0), ga + be.act_bwd(x.contiguous(), gy, 8, 2.0, 0))
