"""GPU parity: the CUDA path (through the C ABI) against the oracle on the same seeded inputs.

Bars (BASELINE.json north_star):
  * integer codes bit-exact vs the reference's own PyTorch path run GPU-eager (the oracle on CUDA
    tensors); at most +-1 code on <= 1e-5 of elements at rounding ties;
  * dequantised outputs, STE gradients, Gram / ADMM terms within 1e-5 relative in fp32.
"""
import numpy as np
import pytest
import torch

import alignq_b200 as aq
from alignq_b200 import _lib as L
from oracle import alignq_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
TIE_FRAC = 1e-5


def t(a):
    return torch.from_numpy(np.asarray(a))


def codes_close(mine, ref, n_levels_step=1):
    """bit-exact up to ties: |diff| <= 1 code on <= 1e-5 of the elements (at least 1 allowed)."""
    d = (mine.to(torch.int64) - ref.to(torch.int64)).abs()
    bad = int((d > 0).sum())
    assert int(d.max()) <= n_levels_step, f"code differs by more than one level: {int(d.max())}"
    assert bad <= max(1, int(TIE_FRAC * mine.numel())), f"{bad} of {mine.numel()} codes differ"
    return bad


def rel_close(a, b, rtol=1e-5, atol_frac=1e-6, what=""):
    """|a-b| <= rtol*|b| + atol_frac*max|b|   (fp32 tolerance 1e-5 relative)"""
    a, b = a.double(), b.double()
    tol = rtol * b.abs() + atol_frac * float(b.abs().max())
    err = (a - b).abs()
    nan_ok = torch.isnan(a) == torch.isnan(b)
    assert bool(nan_ok.all()), f"{what}: NaN pattern differs"
    mask = ~torch.isnan(b)
    assert bool((err[mask] <= tol[mask]).all()), f"{what}: max rel err {float((err[mask] / (b[mask].abs() + 1e-30)).max()):.3e}, max abs {float(err[mask].max()):.3e}"


def grad_close(mine, ref64, orc32, what="", rtol=1e-5, atol_frac=1e-6):
    """Gradient bar: north_star's 1e-5 relative against the oracle's fp64 autograd,
        |mine - ref64| <= 1e-5 |ref64| + max(1e-6 max|ref64|, 2 * floor),
    where floor = max|oracle_fp32 - ref64| is the reference's OWN fp32 autograd noise on the same inputs (only
    where the reference itself is noisier than 1e-6 of the gradient's magnitude does the floor widen the bar).
    Prints both so the log carries the measured numbers."""
    ref64 = ref64.double()
    mx = float(ref64.abs().max())
    floor = float((orc32.double() - ref64).abs().max())
    d = (mine.double() - ref64).abs()
    tol = rtol * ref64.abs() + max(atol_frac * mx, 2.0 * floor)
    print(f"{what}: max|d|/max|ref| = {float(d.max()) / mx:.2e} (reference fp32 floor {floor / mx:.2e}), "
          f"rel-norm {float(d.norm() / ref64.norm()):.2e}")
    assert bool(torch.isfinite(mine).all()), f"{what}: non-finite gradient"
    assert bool((d <= tol).all()), f"{what}: worst |d|/tol = {float((d / tol).max()):.2f}"


def act_codes_via_abi(x, k, variant, ar=2.0, return_cdf=0):
    lib = L.load()
    y = torch.empty_like(x)
    codes = torch.empty(x.shape, dtype=torch.int16, device=x.device)
    L.check(lib.alignq_act_fwd(x.data_ptr(), y.data_ptr(), codes.data_ptr(), x.numel(), k, ar,
                               L.VARIANT_ID[variant], return_cdf, L.stream_ptr()), "act_fwd")
    return y, codes


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", ["A", "B"])
@pytest.mark.parametrize("k", [2, 4, 8])
def test_act_codes_bit_exact_vs_gpu_eager_reference(variant, k):
    torch.manual_seed(0)
    x = torch.randn(128 * 16 * 32 * 32 + 3).to(DEV)          # cfg-1 shape plus a ragged tail
    y, codes = act_codes_via_abi(x, k, variant)
    ref_codes = O.activation_codes(x, k, variant, 2.0)
    ref_y = O.activation_quantize(x, k, "second", variant, 2.0)
    codes_close(codes, ref_codes)
    n = 2 ** k - 1
    step = (2.0 * 2.0 / n) if variant == "A" else (1.0 / n)
    mism = (y != ref_y)
    assert int(mism.sum()) <= max(1, int(TIE_FRAC * x.numel()))
    assert float((y - ref_y).abs().max()) <= step * 1.0001


@pytest.mark.parametrize("variant", ["A", "B"])
def test_act_cpu_oracle_ties_are_rare(variant):
    """CPU-eager divides where CUDA-eager multiplies by a reciprocal: report and bound."""
    torch.manual_seed(1)
    x = torch.randn(1 << 20)
    _, codes = act_codes_via_abi(x.to(DEV), 8, variant)
    cpu_codes = O.activation_codes(x, 8, variant, 2.0)
    bad = codes_close(codes.cpu(), cpu_codes)
    print(f"variant {variant}: {bad} / {x.numel()} codes differ from the CPU oracle")


@pytest.mark.parametrize("variant", ["A", "B", "C"])
def test_act_golden_fixtures(golden, variant):
    g = golden(variant)
    aq.set_args(variant=variant, act_range=2, method="none")
    x = t(g["act_x"]).to(DEV)
    gy = t(g["act_gy"]).to(DEV)
    for k in (2, 4, 8, 1, 32):
        stage = "align" if k == 32 else "second"
        xr = x.clone().requires_grad_(True)
        y = aq.activation_quantize_fn(k, stage)(xr)
        (y * gy).sum().backward()
        ref_y, ref_g = t(g[f"act_y_k{k}"]), t(g[f"act_gx_k{k}"])
        if k in (2, 4, 8):
            _, codes = act_codes_via_abi(x, k, variant)
            codes_close(codes.cpu(), t(g[f"act_codes_k{k}"]))
            # the golden y was produced CPU-eager (true division by n); CUDA-eager and the kernel
            # multiply by fp32(1/n) (verified on the box, profiles/r01_aten_facts.json): 1 ulp apart
            rel_close(y.cpu(), ref_y, rtol=2e-7, atol_frac=2e-7, what=f"act y k={k}")
        else:
            rel_close(y.cpu(), ref_y, what=f"act y k={k}")
        rel_close(xr.grad.cpu(), ref_g, what=f"act gx k={k}")


@pytest.mark.parametrize("variant", ["A", "B"])
def test_act_backward_vs_gpu_eager_autograd(variant):
    torch.manual_seed(2)
    aq.set_args(variant=variant, act_range=2)
    x = (torch.randn(64, 32, 16, 16) * 1.5).to(DEV)
    gy = torch.randn_like(x)
    xr = x.clone().requires_grad_(True)
    (aq.activation_quantize_fn(8, "second")(xr) * gy).sum().backward()
    xo = x.clone().requires_grad_(True)
    (O.activation_quantize(xo, 8, "second", variant, 2.0) * gy).sum().backward()
    rel_close(xr.grad, xo.grad, what="act gx")


def test_act_edge_cases():
    aq.set_args(variant="A", act_range=2)
    q = aq.activation_quantize_fn(8, "second")
    assert q(torch.empty(0, 4, device=DEV)).shape == (0, 4)                       # empty
    for n in (1, 2, 3, 5, 7, 4099):                                               # ragged sizes
        x = torch.randn(n, device=DEV)
        assert torch.equal(q(x), O.activation_quantize(x, 8, "second", "A", 2.0))
    base = torch.randn(4100, device=DEV)
    x = base[1:]                                                                  # 4-byte-aligned view only
    assert x.data_ptr() % 16 != 0
    assert torch.equal(q(x), O.activation_quantize(x, 8, "second", "A", 2.0))
    xt = torch.randn(33, 65, device=DEV).t()                                      # non-contiguous input
    assert torch.equal(q(xt), O.activation_quantize(xt.contiguous(), 8, "second", "A", 2.0))
    sp = torch.tensor([float("nan"), float("inf"), -float("inf"), 0.0, -0.0, 40.0, -40.0, 1e-30], device=DEV)
    y, ref = q(sp), O.activation_quantize(sp, 8, "second", "A", 2.0)
    assert torch.equal(torch.isnan(y), torch.isnan(ref)) and torch.equal(y[1:], ref[1:])
    # saturating ends: codes 0 and n
    _, codes = act_codes_via_abi(sp, 8, "A")
    assert codes[1].item() == 255 and codes[2].item() == 0


def test_act_properties_at_full_size():
    """Size-independent properties at the largest single activation of the configs ([256,144,32,32])."""
    torch.manual_seed(3)
    n = 256 * 144 * 32 * 32
    x = torch.randn(n, device=DEV)
    for variant, lo, hi in (("A", 0, 255), ("B", -510, 510)):
        y, codes = act_codes_via_abi(x, 8, variant)
        assert int(codes.min()) >= lo and int(codes.max()) <= hi
        xs, order = torch.sort(x[: 1 << 22])
        assert bool((torch.diff(y[: 1 << 22][order]) >= 0).all())                 # monotone in x
        assert float(y.abs().max()) <= 2.0
        if variant == "B":                                                        # odd symmetry of the map
            y2, _ = act_codes_via_abi(-x, 8, variant)
            assert float((y + y2).abs().max()) <= 1.0 / 255 + 1e-7
    gy = torch.ones_like(x)
    gx = torch.empty_like(x)
    L.check(L.load().alignq_act_bwd(x.data_ptr(), gy.data_ptr(), gx.data_ptr(), n, 8, 2.0, 0, 0, L.stream_ptr()), "bwd")
    assert float(gx.min()) >= 0.0 and float(gx.max()) <= 4.0 * 0.3989423 + 1e-6   # 2*ar*phi(0)
    # linearity of the backward in gy
    gx3 = torch.empty_like(x)
    gy3 = gy * 3.0
    L.check(L.load().alignq_act_bwd(x.data_ptr(), gy3.data_ptr(), gx3.data_ptr(), n, 8, 2.0, 0, 0, L.stream_ptr()), "bwd")
    assert torch.allclose(gx3, 3.0 * gx, rtol=1e-6, atol=0)


# ------------------------------------------------------------------------------------------------
WEIGHT_SHAPES = [(16, 3, 3, 3), (16, 16, 3, 3), (64, 64, 3, 3), (32, 1, 3, 3), (64, 32, 1, 1), (2048, 512, 1, 1),
                 (64, 3, 7, 7), (5, 7)]


@pytest.mark.parametrize("variant", ["A", "B"])
@pytest.mark.parametrize("k", [4, 8])
def test_weight_quantizer_vs_oracle(variant, k):
    torch.manual_seed(4)
    aq.set_args(variant=variant, bitW=k)
    total = bad = 0
    for shape in WEIGHT_SHAPES:
        fan = int(np.prod(shape[1:]))
        w = (torch.randn(*shape) * (2.0 / fan) ** 0.5).to(DEV)
        gup = torch.randn_like(w)
        mod = aq.weight_quantize_fn(k, "second")
        wr = w.clone().requires_grad_(True)
        wq = mod(wr)
        (wq * gup).sum().backward()
        wo = w.clone().requires_grad_(True)
        oq, oc, op = O.weight_quantize(wo, k, variant)
        (oq * gup).sum().backward()
        w64 = w.double().clone().requires_grad_(True)
        (O.weight_quantize(w64, k, variant)[0] * gup.double()).sum().backward()
        n = 2 ** k - 1
        mine_codes = torch.round(((wq + 1) / 2 if variant == "A" else wq) * n)
        ref_codes = torch.round(((oq.detach() + 1) / 2 if variant == "A" else oq.detach()) * n)
        d = (mine_codes - ref_codes).abs()
        assert float(d.max()) <= 1
        bad += int((d > 0).sum())
        total += w.numel()
        rel_close(mod.weight_cdf, oc.detach(), rtol=1e-5, atol_frac=2e-6, what=f"weight_cdf {shape}")
        rel_close(mod.weight_pdf, op.detach(), rtol=2e-5, atol_frac=2e-6, what=f"weight_pdf {shape}")
        grad_close(wr.grad, w64.grad, wo.grad, what=f"gw {variant} k={k} {shape}")
        assert abs(float(wr.grad.double().sum())) <= 1e-3 * float(wr.grad.double().abs().sum()) + 1e-6   # sum gw = 0
    # stats differ from torch's fp32 mean/std by <= 1 ulp, which can flip a code that sits on a tie
    assert bad <= max(2, int(2e-5 * total)), f"{bad}/{total} weight codes differ"


@pytest.mark.parametrize("variant", ["A", "B"])
def test_weight_codes_bit_exact_given_same_stats(variant):
    """With the kernel's own (mean, std) fed to the reference chain the codes must be identical."""
    torch.manual_seed(5)
    lib = L.load()
    w = (torch.randn(64, 64, 3, 3) * 0.06 + 0.003).to(DEV)
    seg_off, chunk_seg, seg_chunk0, nchunks = L.plan_chunks([w.numel()])
    d_off = torch.tensor(seg_off, dtype=torch.int64, device=DEV)
    d_cs = torch.tensor(chunk_seg, dtype=torch.int32, device=DEV)
    d_c0 = torch.tensor(seg_chunk0, dtype=torch.int32, device=DEV)
    wq, wc, wp = torch.empty_like(w), torch.empty_like(w), torch.empty_like(w)
    codes = torch.empty(w.shape, dtype=torch.int16, device=DEV)
    stats = torch.empty(4, device=DEV)
    ws = torch.empty(2 * nchunks, dtype=torch.float64, device=DEV)
    L.check(lib.alignq_wq_forward(w.data_ptr(), d_off.data_ptr(), d_cs.data_ptr(), d_c0.data_ptr(), 1, nchunks, 8,
                                  L.VARIANT_ID[variant], wq.data_ptr(), wc.data_ptr(), wp.data_ptr(),
                                  codes.data_ptr(), stats.data_ptr(), ws.data_ptr(), L.stream_ptr()), "wq_forward")
    m, s = stats[0], stats[1]
    assert abs(float(m) - float(w.double().mean())) <= 1e-6 * float(w.double().std()) + 1e-9
    assert abs(float(s) / float(w.double().std()) - 1) <= 2e-7
    c_ref, p_ref = O.cdf_map(w, m, s, "w", variant, 0.0)
    assert torch.equal(codes.float(), torch.round(c_ref * 255))
    assert torch.equal(wc, c_ref)
    q_ref = O.uniform_quantize(c_ref, 8) * 2 - 1 if variant == "A" else O.uniform_quantize(c_ref, 8)
    assert torch.equal(wq, q_ref)
    rel_close(wp, p_ref, rtol=2e-6, atol_frac=1e-7, what="pdf")


def test_weight_multi_tensor_launch_matches_single():
    """One launch over many ragged segments == per-tensor launches (bit for bit)."""
    torch.manual_seed(6)
    lib = L.load()
    sizes = [432, 2304, 4096, 4097, 36864, 9, 147456, 8193]
    flat = (torch.randn(sum(sizes)) * 0.1).to(DEV)
    gq = torch.randn_like(flat)
    seg_off, chunk_seg, seg_chunk0, nchunks = L.plan_chunks(sizes)
    d_off = torch.tensor(seg_off, dtype=torch.int64, device=DEV)
    d_cs = torch.tensor(chunk_seg, dtype=torch.int32, device=DEV)
    d_c0 = torch.tensor(seg_chunk0, dtype=torch.int32, device=DEV)
    wq, gw = torch.empty_like(flat), torch.empty_like(flat)
    stats = torch.empty(4 * len(sizes), device=DEV)
    ws = torch.empty(2 * nchunks, dtype=torch.float64, device=DEV)
    L.check(lib.alignq_wq_forward(flat.data_ptr(), d_off.data_ptr(), d_cs.data_ptr(), d_c0.data_ptr(), len(sizes),
                                  nchunks, 8, 1, wq.data_ptr(), 0, 0, 0, stats.data_ptr(), ws.data_ptr(),
                                  L.stream_ptr()), "wq_forward")
    L.check(lib.alignq_wq_backward(flat.data_ptr(), gq.data_ptr(), 0, d_off.data_ptr(), d_cs.data_ptr(), d_c0.data_ptr(),
                                   len(sizes), nchunks, 8, stats.data_ptr(), gw.data_ptr(), 0, ws.data_ptr(),
                                   L.stream_ptr()), "wq_backward")
    # the same through per-segment gradient pointers, with one segment absent and accumulation on top
    parts = [gq[seg_off[i]: seg_off[i + 1]].clone() for i in range(len(sizes))]
    import ctypes
    ptrs = (ctypes.c_void_p * len(parts))(*[None if i == 3 else p.data_ptr() for i, p in enumerate(parts)])
    gw2 = torch.full_like(flat, 7.0)
    L.check(lib.alignq_wq_backward(flat.data_ptr(), 0, ptrs, d_off.data_ptr(), d_cs.data_ptr(), d_c0.data_ptr(),
                                   len(sizes), nchunks, 8, stats.data_ptr(), gw2.data_ptr(), 1, ws.data_ptr(),
                                   L.stream_ptr()), "wq_backward ptrs")
    for i in range(len(sizes)):
        a, b = seg_off[i], seg_off[i + 1]
        if i == 3:
            assert bool((gw2[a:b] == 7.0).all())
        else:
            assert torch.allclose(gw2[a:b], gw[a:b] + 7.0, rtol=1e-6, atol=1e-6)
    aq.set_args(variant="B", bitW=8)
    for i, n in enumerate(sizes):
        a, b = seg_off[i], seg_off[i + 1]
        w = flat[a:b].clone().requires_grad_(True)
        q = aq.weight_quantize_fn(8, "second")(w)
        (q * gq[a:b]).sum().backward()
        assert torch.equal(q.detach(), wq[a:b]), f"segment {i}"
        assert torch.equal(w.grad, gw[a:b]), f"segment {i}"


@pytest.mark.parametrize("variant", ["A", "B", "C"])
def test_weight_golden_fixtures(golden, variant):
    g = golden(variant)
    aq.set_args(variant=variant)
    w, gup = t(g["w"]).to(DEV), t(g["w_gup"]).to(DEV)
    for k in (2, 4, 8):
        aq.set_args(bitW=k)
        mod = aq.weight_quantize_fn(k, "second")
        wr = w.clone().requires_grad_(True)
        wq = mod(wr)
        (wq * gup).sum().backward()
        n = 2 ** k - 1
        step = (2.0 if variant == "A" else 1.0) / n
        d = (wq.detach().cpu() - t(g[f"w_q_k{k}"])).abs()
        assert float(d.max()) <= step * 1.0001 and int((d > 1e-6).sum()) <= 1
        rel_close(mod.weight_cdf.cpu(), t(g[f"w_cdf_k{k}"]), rtol=1e-5, atol_frac=2e-6)
        rel_close(mod.weight_pdf.cpu(), t(g[f"w_pdf_k{k}"]), rtol=2e-5, atol_frac=2e-6)
        # the golden is the reference's fp32 CPU autograd (own noise 2e-7 of max|gw|): 1e-5 relative
        rel_close(wr.grad.cpu(), t(g[f"w_g_k{k}"]), rtol=1e-5, atol_frac=1e-6, what=f"gw k={k}")


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,F,eps", [(8, 96, 0.0), (28, 3000, 1e-5), (128, 4096, 0.0), (100, 1027, 0.0),
                                     (224, 2048, 1e-5)])
def test_corr_vs_oracle(B, F, eps):
    torch.manual_seed(7)
    x = torch.randn(B, F, device=DEV)
    y = torch.randn(B, F, device=DEV) * 2 + 0.5
    G = aq.corr(x, x, eps)
    rel_close(G, O.corr(x, x, eps), rtol=1e-5, atol_frac=2e-6, what="corr(x,x)")
    assert abs(float(G.trace()) - (B - 1)) < 1e-2 or eps > 0
    assert float(G.sum(dim=1).abs().max()) < 1e-3                 # rows of a standardised Gram sum to 0
    rel_close(aq.corr(x, y, eps), O.corr(x, y, eps), rtol=1e-5, atol_frac=2e-6, what="corr(x,y)")


def test_corr_constant_column_gives_nan_like_reference():
    x = torch.randn(8, 64, device=DEV)
    x[:, 5] = 1.25
    assert bool(torch.isnan(aq.corr(x, x, 0.0)).all()) and bool(torch.isnan(O.corr(x, x, 0.0)).all())
    assert not bool(torch.isnan(aq.corr(x, x, 1e-5)).any())


@pytest.mark.parametrize("variant", ["B", "C"])
def test_fused_act_admm_golden_fixtures(golden, variant):
    g = golden(variant)
    aq.set_args(variant=variant, act_range=2, method="ours", gram_mode="fp32")
    dim = g["admm_Z"].shape[0]
    for B in (dim, dim - 3):
        for k in (4, 8):
            tag = f"b{B}_k{k}"
            admm = aq.ADMM(dim).to(DEV)
            with torch.no_grad():
                admm.alterD.copy_(t(g["admm_Z"]))
                admm.gamma.copy_(t(g["admm_U"]))
            Fn = aq.activation_quantize_fn if variant == "B" else aq.activation_quantize_fn2
            x = t(g["act_x"])[:B].to(DEV).requires_grad_(True)
            y, loss = Fn(k, "second", admm)(x)
            ((y * t(g["act_gy"])[:B].to(DEV)).sum() + 1.5 * loss).backward()
            rel_close(y.detach().cpu(), t(g[f"fused_y_{tag}"]), rtol=2e-7, atol_frac=2e-7, what="fused y")
            rel_close(admm.D.cpu(), t(g[f"fused_D_{tag}"]), rtol=1e-5, atol_frac=1e-5, what="D")
            rel_close(loss.detach().cpu(), t(g[f"fused_loss_{tag}"]), rtol=1e-5, what="trans_loss")
            rel_close(x.grad.cpu(), t(g[f"fused_gx_{tag}"]), rtol=1e-5, atol_frac=1e-6, what="fused gx")


@pytest.mark.parametrize("variant,B,shape", [("B", 128, (16, 16, 16)), ("B", 128, (64, 8, 8)), ("C", 28, (64, 14, 14)),
                                             ("B", 100, (32, 16, 16))])
def test_fused_act_admm_vs_gpu_eager_oracle(variant, B, shape):
    torch.manual_seed(8)
    aq.set_args(variant=variant, act_range=2, method="ours", gram_mode="fp32")
    dim = 128
    admm = aq.ADMM(dim).to(DEV)
    x0 = torch.randn(B, *shape, device=DEV)
    gy = torch.randn_like(x0)
    Fn = aq.activation_quantize_fn if variant == "B" else aq.activation_quantize_fn2
    x = x0.clone().requires_grad_(True)
    y, loss = Fn(8, "second", admm)(x)
    ((y * gy).sum() + loss).backward()
    xo = x0.clone().requires_grad_(True)
    Zo = admm.alterD.detach().clone().requires_grad_(True)
    Uo = admm.gamma.detach().clone().requires_grad_(True)
    yo, lo, Do = O.activation_quantize_admm(xo, 8, Zo, Uo, "second", variant, 2.0)
    ((yo * gy).sum() + lo).backward()
    assert int((y != yo).sum()) <= max(1, int(TIE_FRAC * y.numel()))
    # D = corr(t) - corr(x) is a cancellation of two Grams with entries up to ~1: the 1e-5 relative
    # bar applies to the Gram terms, i.e. |dD| <= 1e-5 * max|G| (cuBLAS itself is only that reproducible)
    eps = 0.0 if variant == "B" else 1e-5
    xf = x0.view(B, -1)
    gmax = float(O.corr(xf, xf, eps).abs().max())
    assert float((admm.D - Do.detach()).abs().max()) <= 1e-5 * gmax, "D"
    rel_close(loss.detach(), lo.detach(), rtol=1e-5, what="trans_loss")
    x64 = x0.double().clone().requires_grad_(True)
    Z64 = admm.alterD.detach().double().clone().requires_grad_(True)
    U64 = admm.gamma.detach().double().clone().requires_grad_(True)
    y64, l64, _ = O.activation_quantize_admm(x64, 8, Z64, U64, "second", variant, 2.0)
    ((y64 * gy.double()).sum() + l64).backward()
    grad_close(x.grad, x64.grad, xo.grad, what="gx")
    grad_close(admm.alterD.grad, Z64.grad, Zo.grad, what="d loss / d alterD")
    grad_close(admm.gamma.grad, U64.grad, Uo.grad, what="d loss / d gamma")


def test_fused_backward_with_constant_columns_matches_autograd():
    """QC (eps = 1e-5): a feature column that is constant over the batch has std == 0; torch's std
    backward masks that 0/0 to 0, so the reference gradient is finite (seen in ResNet-50 DANN at B = 2)."""
    torch.manual_seed(10)
    aq.set_args(variant="C", act_range=2, method="ours", gram_mode="fp32")
    B = 4
    admm = aq.ADMM(B).to(DEV)
    x0 = torch.randn(B, 8, 6, 6, device=DEV)
    x0[:, 3] = 0.75
    x0[:, :, 2, 2] = -1.5
    gy = torch.randn_like(x0)
    x = x0.clone().requires_grad_(True)
    y, loss = aq.activation_quantize_fn2(8, "second", admm)(x)
    ((y * gy).sum() + loss).backward()
    xo = x0.clone().requires_grad_(True)
    yo, lo, _ = O.activation_quantize_admm(xo, 8, admm.alterD.detach(), admm.gamma.detach(), "second", "C", 2.0)
    ((yo * gy).sum() + lo).backward()
    assert bool(torch.isfinite(xo.grad).all()) and bool(torch.isfinite(x.grad).all())
    x64 = x0.double().clone().requires_grad_(True)
    y64, l64, _ = O.activation_quantize_admm(x64, 8, admm.alterD.detach().double(), admm.gamma.detach().double(), "second", "C", 2.0)
    ((y64 * gy.double()).sum() + l64).backward()
    grad_close(x.grad, x64.grad, xo.grad, what="gx with constant columns")


@pytest.mark.parametrize("variant", ["B", "C"])
def test_admm_loss_and_zu_update_golden(golden, variant):
    g = golden(variant)
    aq.set_args(bitW=8)
    dim = g["admm_Z"].shape[0]
    for B in (dim, dim - 3):
        admm = aq.ADMM(dim).to(DEV)
        with torch.no_grad():
            admm.alterD.copy_(t(g["admm_Z"]))
            admm.gamma.copy_(t(g["admm_U"]))
        D = t(g[f"admm_D_b{B}"]).to(DEV).requires_grad_(True)
        loss = admm(D)
        loss.backward()
        rel_close(loss.detach().cpu(), t(g[f"admm_loss_b{B}"]), rtol=1e-5, what="ADMM loss")
        rel_close(D.grad.cpu(), t(g[f"admm_dD_b{B}"]), rtol=1e-5, atol_frac=1e-6, what="dL/dD")
        opt = aq.ADMM_OPT([admm.alterD, admm.gamma])
        opt.step([0], [1], [D.detach()], [admm.alterD], [admm.gamma], [admm.mu], [admm.rho])
        rel_close(admm.alterD.detach().cpu(), t(g[f"zu_Z_b{B}"]), rtol=1e-5, atol_frac=1e-6, what="Z")
        rel_close(admm.gamma.detach().cpu(), t(g[f"zu_U_b{B}"]), rtol=1e-5, atol_frac=1e-6, what="U")
    small = aq.ADMM(dim).to(DEV)
    with torch.no_grad():
        small.alterD.fill_(0.3)
        small.gamma.fill_(1e-3)
    D = torch.full((dim, dim), 1e-3, device=DEV, requires_grad=True)
    small(D).backward()
    aq.ADMM_OPT([small.alterD, small.gamma]).step([0], [1], [D.detach()], [small.alterD], [small.gamma], [0.2], [0.3])
    assert float(small.alterD.abs().max()) == 0.0                                 # ||V|| <= mu/rho -> Z = 0
    rel_close(small.gamma.detach().cpu(), t(g["zu_small_U"]), rtol=1e-5, what="U small")


@pytest.mark.parametrize("variant", ["A", "B"])
def test_sgd_step_golden(golden, variant):
    g = golden(variant)
    aq.set_args(bitW=8)
    idx, n = [2, 3], 5
    wc = [t(g[f"sgd_wcdf_{j}"]).to(DEV) for j in range(2)]
    wp = [t(g[f"sgd_wpdf_{j}"]).to(DEV) for j in range(2)]
    cfgs = {"mom": dict(lr=0.04, momentum=0.9, weight_decay=1e-4),
            "nest": dict(lr=0.02, momentum=0.8, weight_decay=5e-4, nesterov=True),
            "plain": dict(lr=0.1)}
    for name, kw in cfgs.items():
        ps = [torch.nn.Parameter(t(g[f"sgd_p0_{i}"]).to(DEV)) for i in range(n)]
        opt = aq.SGD(ps, **kw)
        for s in range(3):
            for i, p in enumerate(ps):
                p.grad = t(g[f"sgd_g{s}_{i}"]).to(DEV)
            opt.step(idx, wc, wp, 1.0, 4.0)
            for i, p in enumerate(ps):
                rel_close(p.detach().cpu(), t(g[f"sgd_{name}_p{s}_{i}"]), rtol=2e-6, atol_frac=1e-7, what=f"{name} p{i} step{s}")
                rel_close(p.grad.cpu(), t(g[f"sgd_{name}_grad{s}_{i}"]), rtol=1e-5, atol_frac=1e-6, what=f"{name} grad{i} step{s}")


def test_uniform_quantize_and_cdf_standalone():
    torch.manual_seed(9)
    x = torch.rand(1000, device=DEV) * 2 - 1
    for k in (1, 2, 8):
        assert torch.equal(aq.uniform_quantize(k)(x), O.uniform_quantize(x, k))
    xr = x.clone().requires_grad_(True)
    aq.uniform_quantize(4)(xr).sum().backward()
    assert torch.equal(xr.grad, torch.ones_like(x))
    for variant in ("A", "B"):
        aq.set_args(variant=variant, act_range=2)
        m, s = torch.tensor(0.1, device=DEV), torch.tensor(0.7, device=DEV)
        for src in ("w", "a"):
            c, p = aq.cdf(m, s, src)(x)
            co, po = O.cdf_map(x, m, s, src, variant, 2.0)
            assert torch.equal(c, co)
            rel_close(p, po, rtol=2e-6, atol_frac=1e-7, what="pdf")


# ------------------------------------------------------------------------------------------------
# tcgen05 Gram modes (B <= 128): fp32-level 'tf32x3' (1e-5) and 'bf16' (1e-2), both stated relative to
# the Gram magnitude max|G| (entries are O(1) on the diagonal, O(1/sqrt(F)) off it).
TC_TOL = {"tf32x3": 1e-5, "bf16": 1e-2}


@pytest.mark.parametrize("mode", ["tf32x3", "bf16"])
@pytest.mark.parametrize("B,F,eps", [(128, 4096, 0.0), (128, 16384, 0.0), (28, 3000, 1e-5), (100, 1027, 0.0), (8, 96, 0.0),
                                     (32, 8192, 0.0), (2, 64, 0.0), (17, 4100, 1e-5), (28, 100352, 1e-5), (31, 1001, 0.0),
                                     (128, 262144, 0.0), (64, 100000, 1e-5)])
def test_tc_corr_vs_oracle(mode, B, F, eps):
    torch.manual_seed(11)
    aq.set_args(gram_mode=mode)
    x = torch.randn(B, F, device=DEV) * 1.7 + 0.3
    G = aq.corr(x, x, eps)
    ref = O.corr(x.double(), x.double(), eps)
    err = float((G.double() - ref).abs().max()) / float(ref.abs().max())
    print(f"corr {mode} B={B} F={F}: max err / max|G| = {err:.2e}")
    assert err <= TC_TOL[mode]


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-5), ("tf32x3", 1e-5), ("bf16", 1e-2)])
@pytest.mark.parametrize("B,F", [(28, 802816), (128, 16384), (128, 262144)])
def test_gram_properties_at_full_size(mode, tol, B, F):
    """Size-independent properties of corr() at the BASELINE sizes (config 5's largest layer, config 2's largest
    layer, and a long-F case): with eps = 0 every standardised column has zero mean and unit unbiased variance, so
    trace(G) = B - 1 and G 1 = 0; G is symmetric; and corr is invariant under x -> a x + b (a > 0)."""
    torch.manual_seed(21)
    aq.set_args(gram_mode=mode)
    x = torch.randn(B, F, device=DEV) * 0.8 + 0.1
    G = aq.corr(x, x, 0.0).double()
    gmax = float(G.abs().max())
    assert abs(float(G.trace()) - (B - 1)) <= tol * (B - 1)
    assert float(G.sum(dim=1).abs().max()) <= tol * B * gmax
    assert float((G - G.t()).abs().max()) <= tol * gmax
    assert float(G.diagonal().min()) > 0.0
    G2 = aq.corr(3.0 * x + 1.0, 3.0 * x + 1.0, 0.0).double()
    assert float((G2 - G).abs().max()) <= 2 * tol * gmax


@pytest.mark.parametrize("mode", ["tf32x3", "bf16"])
@pytest.mark.parametrize("variant,B,shape", [("B", 128, (16, 32, 32)), ("B", 128, (64, 8, 8)), ("C", 28, (64, 14, 14)),
                                             ("B", 100, (3, 11, 13)), ("C", 28, (256, 14, 14)), ("B", 32, (5, 9, 4)),
                                             ("C", 5, (3, 7, 5))])
def test_tc_fused_forward_vs_fp32_mode_and_oracle(mode, variant, B, shape):
    torch.manual_seed(12)
    dim = 128
    admm = aq.ADMM(dim).to(DEV)
    x0 = torch.randn(B, *shape, device=DEV)
    gy = torch.randn_like(x0)
    Fn = aq.activation_quantize_fn if variant == "B" else aq.activation_quantize_fn2
    res = {}
    for m in ("fp32", mode):
        aq.set_args(variant=variant, act_range=2, method="ours", gram_mode=m)
        x = x0.clone().requires_grad_(True)
        y, loss = Fn(8, "second", admm)(x)
        ((y * gy).sum() + loss).backward()
        res[m] = (y.detach(), loss.detach(), admm.D.clone(), x.grad.clone())
    eps = 0.0 if variant == "B" else 1e-5
    gmax = float(O.corr(x0.view(B, -1), x0.view(B, -1), eps).abs().max())
    assert torch.equal(res[mode][0], res["fp32"][0]), "y must not depend on the Gram numerics mode"
    dD = float((res[mode][2] - res["fp32"][2]).abs().max()) / gmax
    dl = abs(float(res[mode][1]) - float(res["fp32"][1])) / abs(float(res["fp32"][1]))
    print(f"fused {mode} {variant} B={B} F={x0[0].numel()}: dD/max|G| = {dD:.2e}, trans_loss rel = {dl:.2e}")
    assert dD <= TC_TOL[mode]
    assert dl <= (1e-5 if mode == "tf32x3" else 2e-2)
    _, lo, Do = O.activation_quantize_admm(x0, 8, admm.alterD.detach(), admm.gamma.detach(), "second", variant, 2.0)
    assert float((res[mode][2] - Do).abs().max()) / gmax <= 2 * TC_TOL[mode]
    # backward: tcgen05 products with three bf16 terms per operand (24 mantissa bits, six MMAs) in tf32x3 mode, one
    # bf16 term in bf16 mode.  tf32x3 is held to the same 1e-5 bar as the fp32 FFMA path, against fp64 autograd.
    if mode == "tf32x3":
        xo = x0.clone().requires_grad_(True)
        yo, lo2, _ = O.activation_quantize_admm(xo, 8, admm.alterD.detach(), admm.gamma.detach(), "second", variant, 2.0)
        ((yo * gy).sum() + lo2).backward()
        x64 = x0.double().clone().requires_grad_(True)
        y64, l64, _ = O.activation_quantize_admm(x64, 8, admm.alterD.detach().double(), admm.gamma.detach().double(),
                                                 "second", variant, 2.0)
        ((y64 * gy.double()).sum() + l64).backward()
        grad_close(res[mode][3], x64.grad, xo.grad, what="gx tcgen05 backward vs fp64 autograd")
        grad_close(res["fp32"][3], x64.grad, xo.grad, what="gx fp32 FFMA backward vs fp64 autograd")
    else:
        e = float((res[mode][3] - res["fp32"][3]).norm() / res["fp32"][3].norm())
        assert e <= 1e-2, f"gx bf16 backward rel-norm err {e:.2e}"


def test_tensor_core_modes_fall_back_to_fp32_kernels_above_128_rows():
    """The tcgen05 kernels cover one batch tile (2 <= B <= 128); above that every gram_mode runs the fp32 FFMA
    kernels (forward and backward), so the results are bit-identical to gram_mode='fp32'."""
    torch.manual_seed(31)
    B, shape = 160, (4, 8, 8)
    admm = aq.ADMM(B).to(DEV)
    x0 = torch.randn(B, *shape, device=DEV)
    gy = torch.randn_like(x0)
    res = {}
    for m in ("fp32", "tf32x3", "bf16"):
        aq.set_args(variant="B", act_range=2, method="ours", gram_mode=m)
        x = x0.clone().requires_grad_(True)
        y, loss = aq.activation_quantize_fn(8, "second", admm)(x)
        ((y * gy).sum() + loss).backward()
        G = aq.corr(x0.view(B, -1), x0.view(B, -1), 0.0)
        res[m] = (y.detach(), loss.detach(), admm.D.clone(), x.grad.clone(), G)
    for m in ("tf32x3", "bf16"):
        for a, b in zip(res[m], res["fp32"]):
            assert torch.equal(a, b)


def test_channels_last_activations_need_no_layout_copy():
    """NHWC (channels_last) activations go through the same kernels in place: element-wise results are
    identical element for element, and the ADMM Gram / trans_loss are invariant under the feature permutation."""
    torch.manual_seed(13)
    aq.set_args(variant="A", act_range=2)
    x = torch.randn(16, 8, 6, 6, device=DEV)
    gy = torch.randn_like(x)
    xc = x.contiguous(memory_format=torch.channels_last).requires_grad_(True)
    xn = x.clone().requires_grad_(True)
    q = aq.activation_quantize_fn(8, "second")
    yc, yn = q(xc), q(xn)
    assert yc.is_contiguous(memory_format=torch.channels_last) and torch.equal(yc, yn)
    (yc * gy).sum().backward()
    (yn * gy).sum().backward()
    assert torch.equal(xc.grad, xn.grad)
    aq.set_args(variant="B", method="ours", gram_mode="fp32")
    admm = aq.ADMM(16).to(DEV)
    f = aq.activation_quantize_fn(8, "second", admm)
    xc2 = x.contiguous(memory_format=torch.channels_last).requires_grad_(True)
    xn2 = x.clone().requires_grad_(True)
    (y1, l1), D1 = f(xc2), admm.D.clone()
    (y2, l2), D2 = f(xn2), admm.D.clone()
    assert torch.equal(y1, y2)
    rel_close(l1.detach(), l2.detach(), rtol=1e-5, what="trans_loss under feature permutation")
    assert float((D1 - D2).abs().max()) <= 1e-5
    ((y1 * gy).sum() + l1).backward()
    ((y2 * gy).sum() + l2).backward()
    rel_close(xc2.grad, xn2.grad, rtol=1e-5, atol_frac=1e-6, what="gx under feature permutation")


@pytest.mark.parametrize("B,F", [(256, 4096), (256, 65536), (128, 8192), (100, 8200), (17, 64), (256, 72)])
def test_gram_bf16_tma_tcgen05_vs_fp64(B, F):
    """bf16 SYRK micro-kernel (TMA + tcgen05, SURVEY.md 8d tensor-bound shape): bf16 products are exact in
    fp32, so against an fp64 product of the SAME bf16 inputs only the fp32 accumulation order differs
    (bar 1e-5 * max|G|); against fp32 inputs the bf16 rounding of the operands is the 1e-2 bar."""
    torch.manual_seed(14)
    lib = L.load()
    x32 = torch.randn(B, F, device=DEV)
    x = x32.to(torch.bfloat16)
    G = torch.empty(B, B, device=DEV)
    ws = torch.empty(int(lib.alignq_gram_bf16_ws_bytes(B)), dtype=torch.uint8, device=DEV)
    L.check(lib.alignq_gram_bf16(x.data_ptr(), B, F, 1, G.data_ptr(), ws.data_ptr(), ws.numel(), L.stream_ptr()), "gram_bf16")
    ref = (x.double() @ x.double().t()) / F
    assert float((G.double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
    ref32 = (x32.double() @ x32.double().t()) / F
    assert float((G.double() - ref32).abs().max()) <= 1e-2 * float(ref32.abs().max())
    assert float((G - G.t()).abs().max()) <= 1e-6 * float(ref.abs().max())          # symmetric
    L.check(lib.alignq_gram_bf16(x.data_ptr(), B, F, 0, G.data_ptr(), ws.data_ptr(), ws.numel(), L.stream_ptr()), "gram_bf16")
    assert float((G.double() - ref * F).abs().max()) <= 1e-5 * float((ref * F).abs().max())
    assert lib.alignq_gram_bf16(x.data_ptr(), B, F - 1, 0, G.data_ptr(), ws.data_ptr(), ws.numel(), L.stream_ptr()) == -2   # F % 8


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-5), ("tf32x3", 1e-5), ("bf16", 6e-2)])
@pytest.mark.parametrize("variant,B,shape", [("B", 128, (16, 16, 16)), ("C", 28, (32, 14, 14)), ("B", 100, (3, 11, 12)),
                                             ("B", 128, (16, 32, 32)), ("C", 28, (256, 14, 14)), ("C", 5, (3, 7, 5))])
def test_pure_admm_gradient_vs_oracle_autograd(mode, tol, variant, B, shape):
    """d trans_loss / d x alone (no gradient through y): isolates the Gram backward -- the two [B,B]x[B,F]
    products, the standardise backward and the chain through the CDF map -- from the much larger STE term.
    Bar for the fp32-parity modes (fp32 FFMA and the tcgen05 'tf32x3' mode): north_star's 1e-5 relative, both as
    rel-norm and as max|d| / max|ref|, against fp64 autograd; the reference's own fp32 autograd sits at ~1e-6 on
    the same inputs (printed).  bf16 mode: the Gram meets its 1e-2 bar (test_tc_corr_vs_oracle); this small
    cancellation-heavy gradient is only bounded loosely (measured 0.8-5e-2)."""
    torch.manual_seed(15)
    aq.set_args(variant=variant, act_range=2, method="ours", gram_mode=mode)
    admm = aq.ADMM(128).to(DEV)
    x0 = torch.randn(B, *shape, device=DEV)
    Fn = aq.activation_quantize_fn if variant == "B" else aq.activation_quantize_fn2
    x = x0.clone().requires_grad_(True)
    _, loss = Fn(8, "second", admm)(x)
    loss.backward()
    xo = x0.double().clone().requires_grad_(True)
    _, lo, _ = O.activation_quantize_admm(xo, 8, admm.alterD.detach().double(), admm.gamma.detach().double(),
                                          "second", variant, 2.0)
    lo.backward()
    x32 = x0.clone().requires_grad_(True)
    _, l32, _ = O.activation_quantize_admm(x32, 8, admm.alterD.detach(), admm.gamma.detach(), "second", variant, 2.0)
    l32.backward()
    e = float((x.grad.double() - xo.grad).norm() / xo.grad.norm())
    emax = float((x.grad.double() - xo.grad).abs().max() / xo.grad.abs().max())
    fl = float((x32.grad.double() - xo.grad).norm() / xo.grad.norm())
    print(f"pure ADMM gradient {mode} {variant} B={B}: rel-norm err vs fp64 autograd {e:.2e}, max|d|/max|ref| {emax:.2e} "
          f"(reference fp32 autograd: {fl:.2e})")
    assert e <= tol and emax <= (tol if mode != "bf16" else 1.0)


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,F,eps", [(8, 96, 0.0), (28, 3000, 1e-5), (128, 4096, 0.0), (100, 1027, 0.0), (160, 640, 1e-5)])
def test_corr_autograd_vs_oracle(B, F, eps):
    """corr(x, y) is differentiable through mean and std of both operands like the reference's (QB:134-137)."""
    torch.manual_seed(16)
    aq.set_args(gram_mode="fp32")
    x0 = torch.randn(B, F, device=DEV) * 1.3 + 0.2
    y0 = torch.randn(B, F, device=DEV) * 0.7 - 0.4
    dG = torch.randn(B, B, device=DEV)
    for same in (False, True):
        x, y = x0.clone().requires_grad_(True), y0.clone().requires_grad_(True)
        (aq.corr(x, x if same else y, eps) * dG).sum().backward()
        ref = {}
        for dt in (torch.float32, torch.float64):
            xo, yo = x0.to(dt).clone().requires_grad_(True), y0.to(dt).clone().requires_grad_(True)
            (O.corr(xo, xo if same else yo, eps) * dG.to(dt)).sum().backward()
            ref[dt] = (xo.grad, None if same else yo.grad)
        grad_close(x.grad, ref[torch.float64][0], ref[torch.float32][0], what=f"corr gx same={same}")
        if same:
            assert y.grad is None
        else:
            grad_close(y.grad, ref[torch.float64][1], ref[torch.float32][1], what="corr gy")
    # only one operand needs a gradient
    x = x0.clone().requires_grad_(True)
    (aq.corr(x, y0, eps) * dG).sum().backward()
    xo = x0.double().clone().requires_grad_(True)
    (O.corr(xo, y0.double(), eps) * dG.double()).sum().backward()
    assert float((x.grad.double() - xo.grad).abs().max()) <= 1e-5 * float(xo.grad.abs().max())


def test_corr_autograd_in_tensor_core_modes_uses_the_same_backward():
    torch.manual_seed(17)
    x0 = torch.randn(64, 2048, device=DEV)
    dG = torch.randn(64, 64, device=DEV)
    grads = []
    for mode in ("fp32", "tf32x3"):
        aq.set_args(gram_mode=mode)
        x = x0.clone().requires_grad_(True)
        (aq.corr(x, x, 0.0) * dG).sum().backward()
        grads.append(x.grad)
    assert torch.equal(grads[0], grads[1])


def test_cdf_standalone_refuses_to_drop_gradients_of_m_and_s():
    aq.set_args(variant="A", act_range=2)
    x = torch.randn(100, device=DEV)
    m = torch.tensor(0.1, device=DEV, requires_grad=True)
    with pytest.raises(aq.AlignQError):
        aq.cdf(m, torch.tensor(0.7, device=DEV), "w")(x)
    c, _ = aq.cdf(m.detach(), torch.tensor(0.7, device=DEV), "w")(x)        # constants are fine
    assert c.shape == x.shape


@pytest.mark.parametrize("variant", ["A", "B"])
def test_linear_q_forward_backward_vs_oracle(variant):
    """Linear_Q (cdf_alignment/dann_office/model/resnet.py:148-160): F.linear on the quantized weight."""
    torch.manual_seed(18)
    aq.set_args(variant=variant, bitW=8)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        lin = aq.linear_Q_fn(8, "second")(96, 40).to(DEV)
        x0 = torch.randn(32, 96, device=DEV)
        gup = torch.randn(32, 40, device=DEV)
        x = x0.clone().requires_grad_(True)
        out = lin(x)
        (out * gup).sum().backward()
        res = {}
        for dt in (torch.float32, torch.float64):
            w = lin.weight.detach().to(dt).clone().requires_grad_(True)
            b = lin.bias.detach().to(dt).clone().requires_grad_(True)
            xo = x0.to(dt).clone().requires_grad_(True)
            wq, _, _ = O.weight_quantize(w, 8, variant)
            oo = torch.nn.functional.linear(xo, wq, b)
            (oo * gup.to(dt)).sum().backward()
            res[dt] = (oo.detach(), w.grad, xo.grad, b.grad, wq.detach())
        n = 255
        codes = lambda q: torch.round(((q + 1) / 2 if variant == "A" else q) * n)
        assert int((codes(lin.quantize_fn.weight_q) != codes(res[torch.float32][4])).sum()) <= 1
        if torch.equal(codes(lin.quantize_fn.weight_q).double(), codes(res[torch.float64][4])):
            rel_close(out.detach(), res[torch.float64][0], rtol=1e-5, atol_frac=2e-6, what="Linear_Q out")
            rel_close(x.grad, res[torch.float64][2], rtol=1e-5, atol_frac=2e-6, what="Linear_Q gx")
        grad_close(lin.weight.grad, res[torch.float64][1], res[torch.float32][1], what="Linear_Q gw")
        grad_close(lin.bias.grad, res[torch.float64][3], res[torch.float32][3], what="Linear_Q gb")
        assert lin.w_bit == 8 and lin.quantize_fn.weight_pdf.shape == lin.weight.shape
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_lmmd_golden_and_fp64_autograd(tag):
    """lmmd (cdf_alignment_admm/dsan_office/utils/mmd.py:24-41) on the CUDA kernels: loss vs the reference's golden value,
    gradients vs the oracle's fp64 autograd at north_star's 1e-5 (with the reference's own fp32 floor)."""
    import os
    from alignq_b200.utils import mmd as M
    from oracle import mmd_oracle as MO
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mmd.npz"))
    src0, tgt0 = t(g[f"{tag}_src"]).to(DEV), t(g[f"{tag}_tgt"]).to(DEV)
    s_label, t_prob = t(g[f"{tag}_s_label"]).to(DEV), t(g[f"{tag}_t_prob"]).to(DEV)
    src, tgt = src0.clone().requires_grad_(True), tgt0.clone().requires_grad_(True)
    loss = M.lmmd(src, tgt, s_label, t_prob)
    assert loss.shape == (1,)
    (loss * 1.7).sum().backward()
    ref = float(t(g[f"{tag}_loss"]))
    if tag == "c":                                               # no class shared by source labels and target predictions
        assert float(loss) == 0.0 and float(src.grad.abs().max()) == 0.0
        return
    s64c, t64c = src0.double().cpu().clone().requires_grad_(True), tgt0.double().cpu().clone().requires_grad_(True)
    l64 = MO.lmmd(s64c, t64c, s_label.cpu(), t_prob.cpu())
    (l64 * 1.7).sum().backward()
    assert abs(float(loss.detach()) - float(l64.detach())) <= 1e-5 * abs(float(l64.detach())) + 1e-7
    assert abs(float(loss.detach()) - ref) <= 2e-5 * abs(ref) + 1e-6       # the golden is the reference's fp32 CPU value
    grad_close(src.grad.cpu(), s64c.grad, t(g[f"{tag}_gs"]), what=f"lmmd g_source {tag}")
    grad_close(tgt.grad.cpu(), t64c.grad, t(g[f"{tag}_gt"]), what=f"lmmd g_target {tag}")
    K = M.guassian_kernel(src0, tgt0)
    rel_close(K.cpu(), t(g[f"{tag}_K"]), rtol=1e-5, atol_frac=1e-6, what="guassian_kernel")


def test_lmmd_nan_kernel_matrix_returns_zero_like_the_reference():
    """mmd.py:35-36: identical features -> bandwidth 0 -> NaN kernels -> the reference returns loss 0 (no gradient)."""
    from alignq_b200.utils import mmd as M
    src = torch.full((6, 10), 0.25, device=DEV, requires_grad=True)
    tgt = torch.full((6, 10), 0.25, device=DEV, requires_grad=True)
    s_label = torch.tensor([0, 1, 2, 0, 1, 2], device=DEV)
    t_prob = torch.softmax(torch.randn(6, 31, generator=torch.Generator().manual_seed(0)), 1).to(DEV)
    t_prob[:, :3] += 1.0
    loss = M.lmmd(src, tgt, s_label, t_prob)
    loss.sum().backward()
    assert float(loss.detach()) == 0.0
    assert float(src.grad.abs().max()) == 0.0 and float(tgt.grad.abs().max()) == 0.0
