"""``SGD`` and ``ADMM_OPT`` -- mirrors of the reference's ``utils/optimizer.py`` (SGD.step:196-262,
ADMM_OPT.step:60-135) with the per-parameter Python/ATen loops replaced by one multi-tensor CUDA
launch (``alignq_sgd_step``) and one launch per ADMM module (``alignq_admm_zu_update``, no host
sync for the soft-threshold branch).

Semantics kept, including the quirks SURVEY.md A.5 lists: ``lam`` / ``lam2`` only rewrite ``p.grad``
(the update uses the momentum buffer, optimizer.py:249-251); the parameter index ``i`` restarts
in every param group while ``idx`` is global; ADMM_OPT skips parameters whose ``.grad`` is None and
pairs every alterD with the gamma that follows it.
Difference: ``p.grad`` is rewritten in place (the reference rebinds ``p.grad.data`` to a new tensor
or to the momentum buffer itself); values are identical.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch.optim.optimizer import Optimizer, required

from .. import _lib as L
from .options import args


class SGD(Optimizer):
    def __init__(self, params, lr=required, momentum=0, dampening=0, weight_decay=0, nesterov=False):
        if lr is not required and lr < 0.0:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if momentum < 0.0:
            raise ValueError("Invalid momentum value: {}".format(momentum))
        if weight_decay < 0.0:
            raise ValueError("Invalid weight_decay value: {}".format(weight_decay))
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov)
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        super().__init__(params, defaults)
        self._table_key = None
        self._table_dev = None
        self._chunk_dev = None
        self._chunk_key = None
        self._pinned = []
        self._table_in_graph_pool = False
        self._table_bound_to_graph = False
        self._deferred = []
        self.defer_uploads_in_capture = False        # set by the owner of a capture that calls flush_deferred_uploads()

    def flush_deferred_uploads(self):
        """After a CUDA-graph capture of ``step``: upload the tables the captured kernels read (see ``step``)."""
        for dst, host in self._deferred:
            dst.copy_(host, non_blocking=True)
        self._deferred = []

    def __setstate__(self, state):
        super().__setstate__(state)
        for group in self.param_groups:
            group.setdefault("nesterov", False)
        self._table_key = None

    def state_dict(self):
        """The reference's checkpoint format (``'optimizer_t': optimizer_t.state_dict()``, main.py:143): per-parameter
        state is exactly ``{'momentum_buffer': tensor}``.  The bookkeeping flag of this implementation is dropped,
        and a buffer that was allocated but never written (no step yet) is not saved at all."""
        sd = super().state_dict()
        state = {}
        for k, st in sd["state"].items():
            if st.get("alignq_first", False):
                continue
            state[k] = {n: v for n, v in st.items() if n != "alignq_first"}
        sd["state"] = state
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        for st in self.state.values():                     # buffers from a checkpoint are initialised
            if "momentum_buffer" in st:
                st["alignq_first"] = False
                st["momentum_buffer"] = st["momentum_buffer"].contiguous() if not L.is_dense(st["momentum_buffer"]) else st["momentum_buffer"]
        self._table_key = None

    def _entries(self, idx, w_cdf, w_pdf):
        idx = list(idx) if idx is not None else []
        use_sur = args.bitW < 32
        ents = []
        for group in self.param_groups:
            wd, mom, damp, nest, lr = (group["weight_decay"], group["momentum"], group["dampening"],
                                       group["nesterov"], group["lr"])
            for i, p in enumerate(group["params"]):
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not L.is_dense(p):
                    raise L.AlignQError("alignq_b200.SGD needs dense fp32 CUDA parameters (no CPU fallback)")
                g = p.grad
                if g.stride() != p.stride():               # the kernel is element-wise over physical memory
                    p.grad = g = torch.empty_like(p).copy_(g)
                buf, first = None, 0
                if mom != 0:
                    st = self.state[p]
                    if "momentum_buffer" not in st:
                        st["momentum_buffer"] = torch.empty_like(p)          # same layout as p
                        st["alignq_first"] = True
                    buf = st["momentum_buffer"]
                    first = 1 if st.get("alignq_first", False) else 0
                cdf_t = pdf_t = None
                if use_sur and i in idx:
                    j = idx.index(i)
                    cdf_t, pdf_t = w_cdf[j].detach(), w_pdf[j].detach()
                    if cdf_t.shape == p.shape and cdf_t.stride() != p.stride():
                        cdf_t = torch.empty_like(p).copy_(cdf_t)
                    if pdf_t.shape == p.shape and pdf_t.stride() != p.stride():
                        pdf_t = torch.empty_like(p).copy_(pdf_t)
                    cdf_t = L.dev_f32_dense(cdf_t, "w_cdf")
                    pdf_t = L.dev_f32_dense(pdf_t, "w_pdf")
                    if cdf_t.numel() != p.numel() or pdf_t.numel() != p.numel():
                        raise L.AlignQError(f"w_cdf/w_pdf #{j} do not match parameter #{i} ({tuple(p.shape)})")
                ents.append((p, g, buf, cdf_t, pdf_t, float(lr), float(mom), float(damp), float(wd), int(nest), first))
        return ents

    @torch.no_grad()
    def step(self, idx=(), w_cdf=(), w_pdf=(), lam=1.0, lam2=4.0, closure=None, grad_scale=1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        ents = self._entries(idx, w_cdf, w_pdf)
        if not ents:
            return loss
        dev = ents[0][0].device
        key = tuple((e[0].data_ptr(), e[1].data_ptr(), L.ptr(e[2]), L.ptr(e[3]), L.ptr(e[4]), e[0].numel()) + e[5:]
                    for e in ents)
        if key != self._table_key:
            arr = (L.SgdTensor * len(ents))()
            for t, e in zip(arr, ents):
                t.p, t.g, t.buf, t.w_cdf, t.w_pdf = e[0].data_ptr(), e[1].data_ptr(), L.ptr(e[2]) or None, \
                    L.ptr(e[3]) or None, L.ptr(e[4]) or None
                t.numel = e[0].numel()
                t.lr, t.momentum, t.dampening, t.weight_decay, t.nesterov, t.first_step = e[5:]
            # pinned staging + async copy: legal inside CUDA-graph capture (becomes a memcpy node that
            # re-reads this pinned buffer on replay, so the buffer is kept alive and never rewritten)
            capturing = dev.type == "cuda" and torch.cuda.is_current_stream_capturing()
            raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            if capturing:
                host = raw.pin_memory()
                self._pinned.append(host)                  # re-read by the graph's memcpy node: lives as long as the graph
            else:
                # eager: a ring of four pinned buffers, each reused once the copy that read it has completed -- the table
                # changes whenever a gradient lands at a new address, which in eager mode can be every step, and one fresh
                # pinned allocation per step (never freed) was a leak and a cudaHostAlloc on the critical path
                ring = self.__dict__.setdefault("_pinned_ring", [])
                k = self.__dict__.get("_pinned_next", 0) % 4
                if len(ring) <= k:
                    ring.append([None, None])
                slot = ring[k]
                if slot[1] is not None:
                    slot[1].synchronize()
                if slot[0] is None or slot[0].numel() != raw.numel():
                    slot[0] = torch.empty(raw.numel(), dtype=torch.uint8).pin_memory()
                slot[0].copy_(raw)
                host = slot[0]
                self._pinned_next = k + 1
            if capturing and self.defer_uploads_in_capture and self._table_dev is not None \
                    and self._table_dev.numel() == host.numel() and not self._table_in_graph_pool:
                # Captured iteration after eager warm-up: the pointers this table holds are fixed for the life of the
                # graph, so the upload does not belong IN the graph (an H2D node costs ~15 us of PCIe latency between the
                # last backward kernel and the update on every replay).  The warm-up's device buffer (ordinary pool,
                # never recycled into the graph's activations) is kept, and the owner of the capture uploads the new
                # contents once, right after the capture ends: flush_deferred_uploads().
                self._deferred.append((self._table_dev, host))
                self._table_bound_to_graph = True
            else:
                if self._table_dev is None or self._table_dev.numel() != host.numel() or self._table_bound_to_graph \
                        or self._table_in_graph_pool:
                    self._table_dev = torch.empty(host.numel(), dtype=torch.uint8, device=dev)
                    self._table_in_graph_pool = capturing
                    self._table_bound_to_graph = False
                self._table_dev.copy_(host, non_blocking=True)     # under capture: a memcpy node re-reading `host`
                if not capturing:
                    ev = torch.cuda.Event()
                    ev.record(torch.cuda.current_stream(dev))
                    self._pinned_ring[(self._pinned_next - 1) % 4][1] = ev
            if self._chunk_dev is None or self._chunk_key != tuple(e[0].numel() for e in ents):
                _, chunk_t, t_chunk0, nchunks = L.plan_chunks([e[0].numel() for e in ents])
                self._chunk_key = tuple(e[0].numel() for e in ents)
                self._chunk_dev = (torch.tensor(chunk_t if nchunks else [0], dtype=torch.int32, device=dev),
                                   torch.tensor(t_chunk0, dtype=torch.int32, device=dev), nchunks)
            self._table_key = key
        chunk_t, t_chunk0, nchunks = self._chunk_dev
        with torch.cuda.device(dev):
            L.check(L.load().alignq_sgd_step(self._table_dev.data_ptr(), chunk_t.data_ptr(), t_chunk0.data_ptr(),
                                             len(ents), nchunks, float(lam), float(lam2), int(min(args.bitW, 32)),
                                             float(grad_scale), L.stream_ptr()), "alignq_sgd_step")
        for e in ents:                               # momentum buffers are initialised now
            if e[10]:
                self.state[e[0]]["alignq_first"] = False
        return loss


class ADMM_OPT(Optimizer):
    def __init__(self, params):
        super().__init__(params, dict())

    @torch.no_grad()
    def step(self, alterD_idx, gamma_idx, Ds, alterDs, gammas, mus, rhos, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = L.load()
        alterD_idx, gamma_idx = list(alterD_idx), list(gamma_idx)
        for group in self.param_groups:
            updated = None                           # (gamma tensor) paired with the last alterD visited
            for i, p in enumerate(group["params"]):
                # the reference skips parameters without .grad (optimizer.py:77-78), i.e. modules that did not take
                # part in the forward.  Z/U whose autograd gradient is deliberately not computed (ADMM.param_grads =
                # False: the closed-form update never reads it) count as present when the caller lists them.
                if p.grad is None and not (getattr(p, "_alignq_closed_form", False)
                                           and (i in alterD_idx or i in gamma_idx)):
                    continue
                if args.bitW < 32 and i in alterD_idx:
                    j = alterD_idx.index(i)
                    D = L.dev_f32(Ds[j].detach(), "D")
                    U = gammas[j].data
                    Z = p.data
                    if not (Z.is_cuda and Z.is_contiguous() and U.is_contiguous() and Z.dtype == torch.float32):
                        raise L.AlignQError("ADMM_OPT needs contiguous fp32 CUDA alterD/gamma (no CPU fallback)")
                    dim, B = Z.shape[0], D.shape[0]
                    with torch.cuda.device_of(Z):
                        L.check(lib.alignq_admm_zu_update(Z.data_ptr(), U.data_ptr(), D.data_ptr(), B, dim, 1,
                                                          float(mus[j]), float(rhos[j]), L.stream_ptr()),
                                "alignq_admm_zu_update")
                    updated = U
                elif args.bitW < 32 and i in gamma_idx:
                    # U <- U + rho (D_ - Z) was applied together with its alterD (optimizer.py:116-124
                    # reuses the D_/alterD locals of the preceding iteration, i.e. exactly that pair).
                    if updated is None or updated.data_ptr() != p.data.data_ptr():
                        raise L.AlignQError("ADMM_OPT: gamma parameter is not preceded by its alterD "
                                            "(the reference relies on that order, optimizer.py:97-124)")
                    updated = None
                else:
                    p.data.add_(p.grad.data, alpha=-group["lr"])     # optimizer.py:127-133 (needs 'lr')
        return loss
