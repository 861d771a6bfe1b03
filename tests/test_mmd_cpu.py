"""CPU: the LMMD oracle against the golden vectors of the reference (tests/golden/mmd.npz, oracle/make_mmd_golden.py) and
the device-side label weights of alignq_b200.utils.mmd.cal_weight (torch ops, no host sync) against the reference's numpy
Weight.cal_weight as restated by the oracle."""
import os

import numpy as np
import torch

from alignq_b200.utils import mmd as M
from oracle import mmd_oracle as MO

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mmd.npz"))


def t(a):
    return torch.from_numpy(np.asarray(a))


def test_oracle_reproduces_reference_goldens():
    for tag in ("a", "b", "c"):
        src, tgt = t(G[f"{tag}_src"]).requires_grad_(True), t(G[f"{tag}_tgt"]).requires_grad_(True)
        loss = MO.lmmd(src, tgt, t(G[f"{tag}_s_label"]), t(G[f"{tag}_t_prob"]))
        assert torch.equal(loss.detach(), t(G[f"{tag}_loss"]))
        assert torch.equal(MO.guassian_kernel(src.detach(), tgt.detach()), t(G[f"{tag}_K"]))
        if loss.requires_grad:
            (loss * 1.7).sum().backward()
            assert torch.equal(src.grad, t(G[f"{tag}_gs"])) and torch.equal(tgt.grad, t(G[f"{tag}_gt"]))
    assert float(t(G["c_loss"])) == 0.0                         # no shared class: count == 0 -> loss 0


def test_device_side_label_weights_match_the_reference_weights():
    for tag in ("a", "b", "c"):
        s_label, t_prob = t(G[f"{tag}_s_label"]), t(G[f"{tag}_t_prob"])
        w_ss, w_tt, w_st = M.cal_weight(s_label, t_prob)
        r_ss, r_tt, r_st = MO.cal_weight(s_label, t_prob)
        B = s_label.shape[0]
        for mine, ref in ((w_ss, r_ss), (w_tt, r_tt), (w_st, r_st)):
            ref = np.broadcast_to(ref, (B, B))
            assert np.allclose(mine.numpy(), ref, rtol=1e-6, atol=1e-9)
