"""CPU: model graphs that host the quantized modules -- state_dict keys, shapes and
named_parameters() order identical to the reference's (tests/golden/model_keys_*.json, written by
oracle/make_model_golden.py from the imported reference); oracle models against the golden logits."""
import json
import os

import numpy as np
import pytest
import torch

import alignq_b200 as aq
from alignq_b200.model import dann, densenet, mobilenetV2, resnet
from oracle import models_oracle as MO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

BUILD = {
    "resnet20_A": ("A", 8, lambda: resnet.resnet20_quant(8, 8, "second")),
    "resnet20_B": ("B", 8, lambda: resnet.resnet20_quant(8, 8, "second")),
    "resnet56_B": ("B", 8, lambda: resnet.resnet56_quant(8, 8, "second")),
    "mobilenetv2_A": ("A", 4, lambda: mobilenetV2.mobile_v2(4, 4, "second")),
    "densenet40_A": ("A", 4, lambda: densenet.densenet_40_quant(8, 8, "second")),
    "resnet50dann_C": ("C", 2, lambda: dann.resnet50_dann(8, 8, "second")),
}


@pytest.mark.parametrize("job", list(BUILD))
def test_state_dict_keys_and_parameter_order_match_reference(job):
    variant, batch, ctor = BUILD[job]
    aq.set_args(variant=variant, train_batch_size=batch, bitW=8, abitW=8)
    model = ctor()
    ref = json.load(open(os.path.join(GOLDEN, f"model_keys_{job}.json")))
    mine = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert list(mine.keys()) == list(ref["state_dict"].keys())
    assert mine == ref["state_dict"]
    assert [n for n, _ in model.named_parameters()] == ref["named_parameters"]


@pytest.mark.parametrize("job,units", [("resnet20_A", [3, 3, 3]), ("resnet20_B", [3, 3, 3])])
def test_oracle_model_reproduces_reference_logits(job, units):
    g = np.load(os.path.join(GOLDEN, f"model_{job}.npz"))
    variant = job[-1]
    x, tgt = torch.from_numpy(g["x"]), torch.from_numpy(g["target"])
    m = MO.OracleResNet(units, 8, 8, variant, 2.0, dim=x.shape[0])
    m.load_state_dict(MO.deterministic_fill(m.state_dict(), seed=11))
    m.train()
    out = m(x)
    logits = out[0] if isinstance(out, tuple) else out
    assert torch.equal(logits.detach(), torch.from_numpy(g["logits"]))
    if variant != "A":
        assert torch.equal(out[1].detach(), torch.from_numpy(g["trans_loss"]))
    tr = MO.OracleTrainer(m)
    for _ in range(2):
        tr.step(x, tgt)
    with torch.no_grad():
        out = m(x)
    logits = out[0] if isinstance(out, tuple) else out
    assert torch.equal(logits, torch.from_numpy(g["logits_after_2_steps"]))


def test_oracle_densenet_reproduces_reference_logits_and_two_training_iterations():
    """oracle/models_oracle.OracleDenseNet (the CPU baseline of the densenet40 workload) against the goldens that
    oracle/make_model_golden.py --job densenet40_A wrote from the imported reference (dense-cifar-10), bit for bit."""
    g = np.load(os.path.join(GOLDEN, "model_densenet40_A.npz"))
    x, tgt = torch.from_numpy(g["x"]), torch.from_numpy(g["target"])
    m = MO.OracleDenseNet(8, 8, 2.0)
    m.load_state_dict(MO.deterministic_fill(m.state_dict(), seed=11))
    m.train()
    assert torch.equal(m(x).detach(), torch.from_numpy(g["logits"]))
    tr = MO.OracleTrainer(m)
    for _ in range(2):
        tr.step(x, tgt)
    with torch.no_grad():
        assert torch.equal(m(x), torch.from_numpy(g["logits_after_2_steps"]))


def test_oracle_mobilenetv2_reproduces_reference_logits_and_two_training_iterations():
    """oracle/models_oracle.OracleMobileNetV2 (W4A4; the CPU baseline of the mobilenetv2 workload) against the goldens of
    oracle/make_model_golden.py --job mobilenetv2_A (imported reference, mobilenet-v2-svhn), bit for bit."""
    g = np.load(os.path.join(GOLDEN, "model_mobilenetv2_A.npz"))
    x, tgt = torch.from_numpy(g["x"]), torch.from_numpy(g["target"])
    m = MO.OracleMobileNetV2(4, 4, 2.0)
    m.load_state_dict(MO.deterministic_fill(m.state_dict(), seed=11))
    m.train()
    assert torch.equal(m(x).detach(), torch.from_numpy(g["logits"]))
    tr = MO.OracleTrainer(m, bitW=4)
    for _ in range(2):
        tr.step(x, tgt)
    with torch.no_grad():
        assert torch.equal(m(x), torch.from_numpy(g["logits_after_2_steps"]))


def test_collectors_follow_reference_indexing():
    from alignq_b200.utils.train import collect_sgd_args, quantized_convs
    aq.set_args(variant="A", bitW=8, abitW=8)
    model = resnet.resnet20_quant(8, 8, "second")
    named = [(n, p) for n, p in model.named_parameters() if "alterD" not in n and "gamma" not in n]
    ref_idx = [j for j, (n, _) in enumerate(named) if "conv" in n and "weight" in n][1:]      # main.py:299-304
    for c in quantized_convs(model):                     # pretend a forward stored the attributes
        c.quantize_fn.weight_cdf = c.weight.detach()
        c.quantize_fn.weight_pdf = c.weight.detach()
    idx, w_cdf, w_pdf = collect_sgd_args(model, [p for _, p in named])
    assert idx == ref_idx and len(w_cdf) == len(idx) == 20
    order = [c.weight for l in model.layers for c in (l.conv0, l.conv1, l.skip_conv) if c is not None]
    assert all(a.data_ptr() == b.data_ptr() for a, b in zip(w_cdf, order))
