"""Where the unmodified reference is mounted (the build container: /root/reference), re-run the golden generator:
it imports the reference's own modules, asserts the oracle bit-identical to them on the seeded inputs, and the
arrays it produces must equal the committed tests/golden/l1_*.npz bit for bit.  On the GPU box the reference is
absent and this test is skipped (the goldens stand in for it there)."""
import os
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "cdf_alignment")), reason="reference not mounted")
def test_committed_goldens_are_what_the_reference_produces(tmp_path):
    env = dict(os.environ, ALIGNQ_GOLDEN_OUT=str(tmp_path), CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, os.path.join(REPO, "oracle", "make_golden.py")], capture_output=True, text=True,
                       timeout=900, cwd=REPO, env=env)
    assert p.returncode == 0, (p.stdout + p.stderr)[-3000:]      # includes the generator's oracle == reference asserts
    for v in "ABC":
        new = np.load(os.path.join(str(tmp_path), f"l1_{v}.npz"))
        old = np.load(os.path.join(REPO, "tests", "golden", f"l1_{v}.npz"))
        assert sorted(new.files) == sorted(old.files)
        for k in new.files:
            assert new[k].shape == old[k].shape and np.array_equal(new[k], old[k], equal_nan=True), (v, k)
