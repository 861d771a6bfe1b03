"""Data-parallel ADMM correlation term with GLOBAL-batch semantics (SURVEY.md 8e option 1, BASELINE.json
north_star (3)): the batch is sharded over the P GPUs of one box, but ``corr`` (QB:134-137) standardises every
feature over the WHOLE batch and its Gram is ``[B_global, B_global]`` across samples, so it does not shard over
the batch.  Feature sharding makes it exact:

  forward   y = act-quant(x)                                     local, element-wise, no collective
            all-to-all: rank r receives ALL B_global rows of its F/P feature columns
            partial sums  S_x = sum_f Xs Xs^T,  S_t = sum_f Ts Ts^T   over the slice (exact column statistics)
            ONE all-reduce (sum) of [2, B_g, B_g]  ->  D = S_t/F - S_x/F,  trans_loss = ADMM.forward(D), dL/dD
  backward  the two [B_g,B_g] x [B_g,F/P] products + standardise backward on the slice (no collective),
            all-to-all back to the row owners, plus the local straight-through term gy * 2 ar phi(x).

Every rank ends up with the identical D, loss and dL/dD (so the redundant Z/U updates of ADMM_OPT stay in lockstep,
``dim`` = B_global) and with d trans_loss / d x for ITS rows.  Because trans_loss is the same global scalar on every
rank, a data-parallel gradient MEAN over ranks must see it scaled by P (``QATStep`` does that).

The exchange logic is separated from the local compute (``backend``): the product backend launches the CUDA kernels
through the C ABI and has no CPU path; the world_size-2 gloo test plugs a torch-eager backend written in the test.

Activation mode switch: ``set_args(dp_gram="feature")`` + ``configure(group)``; default ``"replica"`` keeps every
rank's own ``[b, b]`` Gram (= the reference at train_batch_size = b, what config 5 quotes).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .. import _lib as L

_state = {"group": None, "world": 1, "rank": 0}


def configure(group=None):
    """Bind the feature-sharded mode to a process group (default: the world group)."""
    if not dist.is_initialized():
        raise L.AlignQError("dp_gram.configure(): torch.distributed is not initialised")
    _state.update(group=group, world=dist.get_world_size(group), rank=dist.get_rank(group))
    return _state["world"]


def world() -> int:
    return _state["world"]


class Exchange:
    """The two all-to-alls and the all-reduce, on [rows, features] matrices in PHYSICAL memory order."""

    def __init__(self, group=None, world=None):
        self.group = _state["group"] if group is None and world is None else group
        self.world = _state["world"] if world is None else world

    def rows_to_features(self, x2d: torch.Tensor) -> torch.Tensor:
        """[b, F] (my rows, all features) -> [P*b, F/P] (all rows in global order, my feature slice)."""
        P = self.world
        b, F = x2d.shape
        if F % P:
            raise L.AlignQError(f"feature-sharded Gram: F = {F} is not divisible by the {P} ranks")
        send = x2d.view(b, P, F // P).transpose(0, 1).contiguous()          # [P, b, F/P]: chunk r goes to rank r
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        return recv.view(P * b, F // P)                                     # chunk r came from rank r: its b rows

    def features_to_rows(self, g2d: torch.Tensor, b: int) -> torch.Tensor:
        """[P*b, F/P] (all rows, my slice) -> [b, F] (my rows, all features): the inverse exchange."""
        P = self.world
        Fs = g2d.shape[1]
        send = g2d.contiguous().view(P, b, Fs)                              # chunk r = rank r's rows
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)                # chunk r = my rows, feature slice r
        return recv.transpose(0, 1).contiguous().view(b, P * Fs)

    def all_reduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t


class CudaBackend:
    """Local compute of the feature-sharded path: the sm_100a kernels behind include/alignq_b200.h."""

    @staticmethod
    def _ws(B, Fdim, device):
        from ..model.quantization import _gram_ws
        return _gram_ws(B, Fdim, device)

    def act_fwd(self, x, a_bit, act_range, variant_id):
        y = torch.empty_like(x)
        with torch.cuda.device_of(x):
            L.check(L.load().alignq_act_fwd(x.data_ptr(), y.data_ptr(), 0, x.numel(), a_bit, act_range, variant_id, 0,
                                            L.stream_ptr()), "alignq_act_fwd")
        return y

    def gram_sums(self, xs, act_range, eps, gram_mode):
        Bg, Fs = xs.shape
        sums = torch.empty(2, Bg, Bg, dtype=torch.float32, device=xs.device)
        ws = self._ws(Bg, Fs, xs.device)
        with torch.cuda.device_of(xs):
            L.check(L.load().alignq_gram_sums_fwd(xs.data_ptr(), Bg, Fs, act_range, eps, sums.data_ptr(), ws.data_ptr(),
                                                  ws.numel(), gram_mode, L.stream_ptr()), "alignq_gram_sums_fwd")
        return sums

    def admm_from_sums(self, sums, F_total, Z, U, mu, rho, d_out=None):
        Bg, dim = sums.shape[1], Z.shape[0]
        dev = sums.device
        D = d_out if (d_out is not None and d_out.shape == (Bg, Bg) and d_out.is_contiguous()) \
            else torch.empty(Bg, Bg, dtype=torch.float32, device=dev)
        dLdD = torch.empty_like(D)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        inv_f = torch.full((1,), 1.0 / float(F_total), dtype=torch.float32, device=dev)
        lib = L.load()
        with torch.cuda.device_of(sums):
            L.check(lib.alignq_gram_sums_to_d(sums.data_ptr(), inv_f.data_ptr(), Bg, 1, D.data_ptr(), L.stream_ptr()),
                    "alignq_gram_sums_to_d")
            L.check(lib.alignq_admm_loss(D.data_ptr(), Bg, Z.data_ptr(), U.data_ptr(), dim, 1, mu, rho, 0, 0,
                                         loss.data_ptr(), dLdD.data_ptr(), 0, 0, L.stream_ptr()), "alignq_admm_loss")
        return D, loss, dLdD

    def slice_bwd(self, xs, dLdD, gloss_over_p, a_bit, act_range, eps, gram_mode):
        """ADMM part of the gradient on the slice (no gy): corr_bwd(x, -g dLdD) + corr_bwd(t, +g dLdD) 2 ar phi(x);
        the kernel divides by the slice's F/P, so the upstream scalar arrives divided by P."""
        Bg, Fs = xs.shape
        g = torch.empty_like(xs)
        ws = self._ws(Bg, Fs, xs.device)
        with torch.cuda.device_of(xs):
            L.check(L.load().alignq_act_admm_bwd(xs.data_ptr(), 0, dLdD.data_ptr(), gloss_over_p.data_ptr(), Bg, Fs, a_bit,
                                                 act_range, eps, g.data_ptr(), ws.data_ptr(), ws.numel(), gram_mode,
                                                 L.stream_ptr()), "alignq_act_admm_bwd (feature slice)")
        return g

    def act_bwd_add(self, x, gy, gadd, a_bit, act_range, variant_id):
        gx = torch.empty_like(x)
        with torch.cuda.device_of(x):
            if gy is None:
                return gadd.view_as(x) if gadd.stride() == x.stride() else gx.copy_(gadd.view(x.shape))
            L.check(L.load().alignq_act_bwd_add(x.data_ptr(), gy.data_ptr(), gadd.data_ptr(), gx.data_ptr(), x.numel(),
                                                a_bit, act_range, variant_id, 0, L.stream_ptr()), "alignq_act_bwd_add")
        return gx

    def act_bwd(self, x, gy, a_bit, act_range, variant_id):
        gx = torch.empty_like(x)
        with torch.cuda.device_of(x):
            L.check(L.load().alignq_act_bwd(x.data_ptr(), gy.data_ptr(), gx.data_ptr(), x.numel(), a_bit, act_range,
                                            variant_id, 0, L.stream_ptr()), "alignq_act_bwd")
        return gx

    def admm_param_grads(self, D, Z, U, mu, rho, gl):
        gZ, gU = torch.empty_like(Z), torch.empty_like(U)
        with torch.cuda.device_of(D):
            L.check(L.load().alignq_admm_loss(D.data_ptr(), D.shape[0], Z.data_ptr(), U.data_ptr(), Z.shape[0], 1, mu, rho,
                                              gl.data_ptr(), 0, 0, 0, gZ.data_ptr(), gU.data_ptr(), L.stream_ptr()),
                    "alignq_admm_loss (parameter grads)")
        return gZ, gU

    @staticmethod
    def prepare(t, what):
        return L.dev_f32_dense(t, what)

    @staticmethod
    def scalar(g):
        return L.dev_f32(g.reshape(1), "grad of trans_loss")


class FeatureShardedAdmmFn(torch.autograd.Function):
    """(y, trans_loss, D) = activation quantizer + GLOBAL-batch ADMM term of a batch-sharded activation."""

    @staticmethod
    def forward(ctx, x, alterD, gamma, a_bit, act_range, eps, mu, rho, gram_mode, variant_id, backend, exch, d_out,
                param_grads):
        xc = backend.prepare(x, "activation")
        Z, U = alterD.detach(), gamma.detach()
        b = xc.shape[0]
        F = xc.numel() // b
        Bg = exch.world * b
        if Z.shape[0] < Bg or Z.shape != U.shape or Z.shape[0] != Z.shape[1]:
            raise L.AlignQError(f"ADMM dim {tuple(Z.shape)} must be square and >= the GLOBAL batch {Bg} "
                                f"(= {exch.world} ranks x {b}) in dp_gram='feature' mode")
        y = backend.act_fwd(xc, a_bit, act_range, variant_id)
        x2d = torch.as_strided(xc, (b, F), (F, 1), xc.storage_offset())       # physical order: any dense layout
        xs = exch.rows_to_features(x2d)
        sums = backend.gram_sums(xs, act_range, eps, gram_mode)
        exch.all_reduce_sum_(sums)
        D, loss, dLdD = backend.admm_from_sums(sums, F, Z, U, mu, rho, d_out)
        ctx.save_for_backward(xc, xs, dLdD, D, Z, U)
        ctx.cfg = (a_bit, act_range, eps, mu, rho, gram_mode, variant_id, backend, exch, bool(param_grads))
        ctx.mark_non_differentiable(D)
        ctx.set_materialize_grads(False)
        return y, loss, D

    @staticmethod
    def backward(ctx, gy, gloss, _gD):
        xc, xs, dLdD, D, Z, U = ctx.saved_tensors
        a_bit, act_range, eps, mu, rho, gram_mode, variant_id, backend, exch, param_grads = ctx.cfg
        none = (None,) * 11
        b = xc.shape[0]
        gx = gZ = gU = None
        if gy is not None and gy.stride() != xc.stride():
            gy = torch.empty_like(xc).copy_(gy)
        if gloss is None:                       # a backward pass that does not reach trans_loss: local STE only
            if gy is not None and ctx.needs_input_grad[0]:
                gx = backend.act_bwd(xc, gy, a_bit, act_range, variant_id)
            return (gx, None, None) + none
        gl = backend.scalar(gloss)
        if ctx.needs_input_grad[0]:
            g_slice = backend.slice_bwd(xs, dLdD, gl / float(exch.world), a_bit, act_range, eps, gram_mode)
            g_rows = exch.features_to_rows(g_slice, b)                        # [b, F], physical order of xc
            gadd = torch.as_strided(g_rows, xc.shape, xc.stride(), g_rows.storage_offset())
            gx = backend.act_bwd_add(xc, gy, gadd, a_bit, act_range, variant_id)
        if param_grads and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
            gZ, gU = backend.admm_param_grads(D, Z, U, mu, rho, gl)
            gZ = gZ if ctx.needs_input_grad[1] else None
            gU = gU if ctx.needs_input_grad[2] else None
        return (gx, gZ, gU) + none


def feature_sharded_act_admm(x, admm, a_bit, act_range, eps, gram_mode_id, variant_id, backend=None, exch=None):
    """Module-level entry used by ``activation_quantize_fn`` when ``args.dp_gram == 'feature'``."""
    backend = CudaBackend() if backend is None else backend
    exch = Exchange() if exch is None else exch
    Bg = exch.world * x.shape[0]
    d_out = admm.d_slot(Bg) if hasattr(admm, "d_slot") else None
    y, loss, D = FeatureShardedAdmmFn.apply(x, admm.alterD, admm.gamma, a_bit, act_range, eps, float(admm.mu), float(admm.rho),
                                            gram_mode_id, variant_id, backend, exch, d_out,
                                            bool(getattr(admm, "param_grads", True)))
    admm.D = D
    return y, loss
