"""Generates tests/golden/mmd.npz from the UNMODIFIED reference (cdf_alignment_admm/dsan_office/utils/{mmd,Weight}.py)
run on CPU in the build container, and asserts that oracle/mmd_oracle.py reproduces it bit for bit.
Shims (outside the reference tree): argparse runs at import (utils/options_office.py) -> a minimal sys.argv; the
module-global `device` of mmd.py (hard-coded cuda:0, mmd.py:7) is pointed at the CPU after import."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/cdf_alignment_admm/dsan_office"
sys.path.insert(0, REPO)


def main():
    sys.argv = ["x", "--bitW", "8", "--abitW", "8"]
    sys.path.insert(0, REF)
    import utils.mmd as ref_mmd                       # noqa: E402  (the reference)
    from utils.Weight import Weight                    # noqa: E402
    ref_mmd.device = torch.device("cpu")
    from oracle import mmd_oracle as MO

    out = {}
    torch.manual_seed(0)
    for tag, B, d, shared in (("a", 8, 32, True), ("b", 28, 256, True), ("c", 6, 16, False)):
        src = torch.randn(B, d) * 0.7
        tgt = torch.randn(B, d) * 0.9 + 0.2
        s_label = torch.randint(0, 31 if shared else 3, (B,))
        logits = torch.randn(B, 31)
        if not shared:
            logits[:, :3] = -50.0                      # the target never predicts the source's classes: count == 0
        t_prob = torch.softmax(logits, dim=1)
        gl = torch.tensor(1.7)
        s1, t1 = src.clone().requires_grad_(True), tgt.clone().requires_grad_(True)
        loss_ref = ref_mmd.lmmd(s1, t1, s_label, t_prob)
        if loss_ref.requires_grad:
            (loss_ref * gl).sum().backward()
        s2, t2 = src.clone().requires_grad_(True), tgt.clone().requires_grad_(True)
        loss_orc = MO.lmmd(s2, t2, s_label, t_prob)
        if loss_orc.requires_grad:
            (loss_orc * gl).sum().backward()
        assert torch.equal(loss_ref, loss_orc), tag
        for a, b in ((s1.grad, s2.grad), (t1.grad, t2.grad)):
            assert (a is None) == (b is None) and (a is None or torch.equal(a, b)), tag
        wr = Weight.cal_weight(s_label, t_prob, type="visual")
        wo = MO.cal_weight(s_label, t_prob)
        assert all(np.array_equal(x, y) for x, y in zip(wr, wo)), tag
        K = ref_mmd.guassian_kernel(src, tgt)
        assert torch.equal(K, MO.guassian_kernel(src, tgt))
        out.update({f"{tag}_src": src.numpy(), f"{tag}_tgt": tgt.numpy(), f"{tag}_s_label": s_label.numpy(),
                    f"{tag}_t_prob": t_prob.numpy(), f"{tag}_loss": loss_ref.detach().numpy(), f"{tag}_K": K.numpy(),
                    f"{tag}_gs": (s1.grad if s1.grad is not None else torch.zeros_like(src)).numpy(),
                    f"{tag}_gt": (t1.grad if t1.grad is not None else torch.zeros_like(tgt)).numpy(),
                    f"{tag}_w_ss": np.broadcast_to(wr[0], (B, B)).copy() if wr[0].shape != (B, B) else wr[0],
                    f"{tag}_w_st": np.broadcast_to(wr[2], (B, B)).copy() if wr[2].shape != (B, B) else wr[2]})
    np.savez_compressed(os.path.join(REPO, "tests", "golden", "mmd.npz"), **out)
    print("wrote tests/golden/mmd.npz:", sorted(out)[:6], "...")


if __name__ == "__main__":
    main()
