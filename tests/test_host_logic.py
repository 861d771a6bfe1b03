"""CPU: host-side mirror of the reference interface -- names, constructor arguments, attributes,
state_dict keys, config handling and the 'no CPU fallback' rule."""
import pytest
import torch
import torch.nn as nn

import alignq_b200 as aq
from alignq_b200.model import quantization as Q


def test_surface_names_match_reference():
    for name in ("uniform_quantize", "cdf", "weight_quantize_fn", "activation_quantize_fn",
                 "activation_quantize_fn2", "corr", "conv2d_Q_fn", "linear_Q_fn"):
        assert hasattr(Q, name)
    for name in ("ADMM", "SGD", "ADMM_OPT", "set_args", "args"):
        assert hasattr(aq, name)


def test_conv2d_q_is_a_conv2d_with_reference_state_dict_keys():
    Conv = aq.conv2d_Q_fn(w_bit=8, stage="second")
    c = Conv(3, 16, kernel_size=3, stride=1, padding=1, bias=False)
    assert isinstance(c, nn.Conv2d) and isinstance(c.quantize_fn, aq.weight_quantize_fn)
    assert list(c.state_dict().keys()) == ["weight"]
    assert c.quantize_fn.w_bit == 8 and c.quantize_fn.stage == "second"
    d = Conv(8, 8, 3, groups=8)                      # depthwise (mobilenetV2.py:40)
    assert sorted(d.state_dict().keys()) == ["bias", "weight"] and d.weight.shape == (8, 1, 3, 3)
    lin = aq.linear_Q_fn(8, "second")(64, 10)
    assert isinstance(lin, nn.Linear) and lin.w_bit == 8


def test_admm_module_attributes_and_state():
    torch.manual_seed(0)
    a = aq.ADMM(16)
    assert a.mu == 0.2 and a.rho == 0.3
    assert a.alterD.shape == (16, 16) and a.gamma.shape == (16, 16)
    assert float(a.alterD.min()) >= 0.0 and float(a.alterD.max()) < 1.0      # torch.rand init
    assert sorted(a.state_dict().keys()) == ["alterD", "gamma"]
    q = aq.activation_quantize_fn(8, "second", a)
    assert q.opt is a and q.a_bit == 8 and q.variant == "B"
    holder = nn.Module()
    holder.admm0, holder.act_q0 = a, q                # same storage under two keys, as in the reference
    assert {"admm0.alterD", "act_q0.opt.alterD"} <= set(holder.state_dict().keys())


def test_identity_paths_do_not_touch_the_gpu():
    x = torch.randn(4, 3)
    assert aq.activation_quantize_fn(32, "second")(x) is x
    y, loss = aq.activation_quantize_fn(32, "second", aq.ADMM(4))(x)
    assert y is x and loss == 0
    w = torch.randn(5, 5)
    wq = aq.weight_quantize_fn(32, "second")
    assert wq(w) is w and wq.weight_cdf is w and wq.weight_q is w
    assert aq.uniform_quantize(32)(x).data_ptr() == x.data_ptr()


def test_no_cpu_fallback():
    x = torch.randn(4, 8)
    with pytest.raises(aq.AlignQError):
        aq.activation_quantize_fn(8, "second")(x)
    with pytest.raises(aq.AlignQError):
        aq.weight_quantize_fn(8, "second")(x)
    with pytest.raises(aq.AlignQError):
        aq.corr(x, x)
    with pytest.raises(aq.AlignQError):
        aq.ADMM(4)(torch.randn(4, 4))
    with pytest.raises(aq.AlignQError):
        aq.uniform_quantize(4)(x)
    p = nn.Parameter(torch.randn(3))
    p.grad = torch.randn(3)
    with pytest.raises(aq.AlignQError):
        aq.SGD([p], lr=0.1).step([], [], [], 1.0, 4.0)


def test_set_args_validation():
    aq.set_args(bitW=8, abitW=8, variant="B", gram_mode="fp32", act_range=2)
    assert aq.args.bitW == 8 and aq.args.variant == "B"
    with pytest.raises(KeyError):
        aq.set_args(bitw=8)
    with pytest.raises(ValueError):
        aq.set_args(variant="Z")
    aq.parse_args(["--bitW", "4", "--abitW", "4", "--act_range", "2", "--lam2", "4"])
    assert aq.args.bitW == 4 and aq.args.lam2 == 4.0


def test_sgd_constructor_validation_like_reference():
    p = nn.Parameter(torch.zeros(2))
    with pytest.raises(ValueError):
        aq.SGD([p], lr=-1.0)
    with pytest.raises(ValueError):
        aq.SGD([p], lr=0.1, momentum=-0.1)
    with pytest.raises(ValueError):
        aq.SGD([p], lr=0.1, nesterov=True)
    o = aq.SGD([p], lr=0.1, momentum=0.9, weight_decay=1e-4)
    assert o.param_groups[0]["momentum"] == 0.9 and o.param_groups[0]["nesterov"] is False
    assert o.step([], [], [], 1.0, 4.0) is None         # no grads -> nothing to do, no GPU needed


def test_product_never_imports_the_oracle():
    import os
    import re
    root = os.path.dirname(os.path.abspath(aq.__file__))
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle[./]alignq_oracle|_ref", re.M)
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not pat.search(src), f"{f} reaches into oracle/"
