"""CPU, world_size = 2 over gloo: the data-parallel GLOBAL-batch ADMM term (alignq_b200/utils/dp_gram.py) --
all-to-all to feature slices, partial Gram sums, one all-reduce, ADMM loss, and the backward with the all-to-all
back -- against the oracle's single-process autograd on the whole batch.

The exchange logic and the autograd Function are the product's; the LOCAL compute plugged in here is a torch-eager
backend written from the oracle (the product's backend launches the CUDA kernels and has no CPU path).  The same
Function with the CUDA backend is checked over NCCL on the GPU box by tests/test_gpu_dp.py."""
import torch
import torch.distributed as dist

from alignq_b200.utils import dp_gram
from oracle import alignq_oracle as O
from oracle import closed_forms as CF
from _dist_util import run2


class TorchBackend:
    """Oracle-based stand-in for dp_gram.CudaBackend (fp64, CPU)."""

    @staticmethod
    def prepare(t, what):
        return t.contiguous()

    @staticmethod
    def scalar(g):
        return g.reshape(1)

    def act_fwd(self, x, a_bit, ar, variant_id):
        return O.activation_quantize(x, a_bit, "second", "B", ar).detach()

    def gram_sums(self, xs, ar, eps, gram_mode):
        t = O.activation_map(xs, "B", ar)

        def S(a):
            s = (a - a.mean(0)) / (a.std(0) + eps)
            return s @ s.t()
        return torch.stack([S(xs), S(t)])

    def admm_from_sums(self, sums, F, Z, U, mu, rho, d_out=None):
        D = sums[1] / F - sums[0] / F
        Dg = D.clone().requires_grad_(True)
        with torch.enable_grad():
            loss = O.admm_loss(Dg, Z, U, mu, rho)
        (dLdD,) = torch.autograd.grad(loss, Dg)
        return D, loss.detach(), dLdD

    def slice_bwd(self, xs, dLdD, gl_over_p, a_bit, ar, eps, gram_mode):
        T = ar * (2.0 * O.normal_cdf(xs, torch.zeros(1, dtype=xs.dtype), torch.ones(1, dtype=xs.dtype)) - 1.0)
        dD = dLdD * gl_over_p
        return CF.corr_backward(xs, -dD, eps) + CF.corr_backward(T, dD, eps) * (2.0 * ar) * CF.phi(xs)

    def act_bwd_add(self, x, gy, gadd, a_bit, ar, variant_id):
        return gadd if gy is None else gadd + gy * (2.0 * ar) * CF.phi(x)

    def act_bwd(self, x, gy, a_bit, ar, variant_id):
        return gy * (2.0 * ar) * CF.phi(x)

    def admm_param_grads(self, D, Z, U, mu, rho, gl):
        Zg, Ug = Z.clone().requires_grad_(True), U.clone().requires_grad_(True)
        with torch.enable_grad():
            loss = O.admm_loss(D, Zg, Ug, mu, rho) * gl.reshape(())
        return torch.autograd.grad(loss, (Zg, Ug))


def _inputs(eps):
    torch.manual_seed(5)
    Bg, shape = 12, (4, 3, 6)                               # F = 72, divisible by 2
    x = torch.randn(Bg, *shape, dtype=torch.float64) * 1.2 + 0.1
    gy = torch.randn(Bg, *shape, dtype=torch.float64)
    Z = torch.rand(16, 16, dtype=torch.float64)             # dim 16 > B_global 12: the [:B, :B] corner is used
    U = torch.rand(16, 16, dtype=torch.float64)
    return x, gy, Z, U


def _single_process_reference(eps, w):
    x, gy, Z, U = _inputs(eps)
    xr = x.clone().requires_grad_(True)
    Zr, Ur = Z.clone().requires_grad_(True), U.clone().requires_grad_(True)
    y, loss, D = O.activation_quantize_admm(xr, 8, Zr, Ur, "second", "B" if eps == 0.0 else "C", 2.0)
    ((y * gy).sum() + w * loss).backward()
    return y.detach(), loss.detach(), D.detach(), xr.grad, Zr.grad, Ur.grad


def _dp_worker(eps, w, ste_only=False):
    def fn(rank, world):
        x, gy, Z, U = _inputs(eps)
        b = x.shape[0] // world
        xl = x[rank * b:(rank + 1) * b].clone().requires_grad_(True)
        gyl = gy[rank * b:(rank + 1) * b]
        Zp, Up = Z.clone().requires_grad_(True), U.clone().requires_grad_(True)
        exch = dp_gram.Exchange(group=None, world=world)
        y, loss, D = dp_gram.FeatureShardedAdmmFn.apply(xl, Zp, Up, 8, 2.0, eps, 0.2, 0.3, 0, 1, TorchBackend(), exch, None, True)
        if ste_only:
            (y * gyl).sum().backward()
        else:
            ((y * gyl).sum() + w * loss).backward()
        return y.detach(), loss.detach(), D, xl.grad, Zp.grad, Up.grad
    return fn


def _dp_full(rank, world):
    return _dp_worker(0.0, 1.7)(rank, world)


def _dp_eps(rank, world):
    return _dp_worker(1e-5, 0.6)(rank, world)


def _dp_ste(rank, world):
    return _dp_worker(0.0, 1.0, ste_only=True)(rank, world)


def _check(out, ref, world=2):
    y, loss, D, gx, gZ, gU = ref
    b = y.shape[0] // world
    for r, (yr, lr, Dr, gxr, gZr, gUr) in enumerate(out):
        assert torch.equal(yr, y[r * b:(r + 1) * b])                                   # element-wise, no collective
        assert torch.allclose(Dr, D, rtol=1e-10, atol=1e-13)                           # identical global D on every rank
        assert abs(float(lr) - float(loss)) <= 1e-12 * abs(float(loss))
        assert torch.allclose(gxr, gx[r * b:(r + 1) * b], rtol=1e-9, atol=1e-13), f"rank {r}: gx"
        assert torch.allclose(gZr, gZ, rtol=1e-10, atol=1e-14) and torch.allclose(gUr, gU, rtol=1e-10, atol=1e-14)
    assert torch.equal(out[0][2], out[1][2])                                           # Z/U updates stay in lockstep


def test_feature_sharded_admm_matches_single_device_autograd():
    _check(run2(_dp_full), _single_process_reference(0.0, 1.7))


def test_feature_sharded_admm_with_eps_variant():
    _check(run2(_dp_eps), _single_process_reference(1e-5, 0.6))


def test_backward_without_trans_loss_is_local_ste_only():
    out = run2(_dp_ste)
    x, gy, Z, U = _inputs(0.0)
    ref = gy * 4.0 * CF.phi(x)
    for r, o in enumerate(out):
        assert torch.allclose(o[3], ref[r * 6:(r + 1) * 6], rtol=1e-12) and o[4] is None


def _exchange_roundtrip(rank, world):
    torch.manual_seed(rank)
    b, F = 3, 8
    x = torch.arange(b * F, dtype=torch.float64).view(b, F) + 100 * rank
    ex = dp_gram.Exchange(group=None, world=world)
    xs = ex.rows_to_features(x)
    back = ex.features_to_rows(xs, b)
    return x, xs, back


def test_exchange_is_a_transpose_of_ownership_and_inverts():
    out = run2(_exchange_roundtrip)
    full = torch.cat([o[0] for o in out])                                              # global [6, 8]
    for r, (x, xs, back) in enumerate(out):
        assert torch.equal(xs, full[:, r * 4:(r + 1) * 4])                             # all rows, my feature slice
        assert torch.equal(back, x)
