"""CPU: the autograd plumbing of the backward fusions in model/fused.py + model/conv_tc.py, with the C-ABI replaced by a
recorder (no kernel runs): which entry points two residual blocks of the C = 16 stage call, in which order.

* `fork`: the shortcut's gradient is parked in the producer's link by the bn-act layer that took it as residual, so no
  autograd accumulate kernel and no zero-filled placeholder appears;
* `fuse_dgrad_bn`: every own data-gradient convolution runs the preceding bn-act layer's reduce pass
  (alignq_conv3x3_bwd_data_bnreduce) and that layer's backward is alignq_bn_act_bwd_apply alone.
The numerics of the same paths are tested on the GPU (tests/test_gpu_conv.py, tests/test_gpu_fused_bn.py)."""
import contextlib
import types

import pytest
import torch

import alignq_b200 as aq
from alignq_b200 import _lib as L
from alignq_b200.model import conv_tc, fused
from alignq_b200.model.resnet import PreActBlock_conv_Q


@pytest.fixture
def recorder(monkeypatch):
    calls = []
    real = L.load()

    class Fake:
        def __getattr__(self, name):
            if name in ("alignq_bn_act_ws_doubles", "alignq_conv3x3_ws_bytes"):
                return getattr(real, name)

            def f(*a):
                calls.append(name)
                return 0
            return f

    fake = Fake()
    monkeypatch.setattr(L, "load", lambda: fake)
    monkeypatch.setattr(L, "stream_ptr", lambda: 0)
    monkeypatch.setattr(L, "like_layout", lambda g, ref, what: g)
    monkeypatch.setattr(fused, "can_fuse", lambda bn, actq, x: True)
    monkeypatch.setattr(conv_tc, "applies", lambda *a, **k: True)
    monkeypatch.setattr(conv_tc, "_workspace", lambda C, dev: torch.empty(16))
    monkeypatch.setattr(conv_tc.WgradStream, "get", classmethod(lambda cls, dev: None))
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: types.SimpleNamespace(cuda_stream=0))
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    yield calls
    aq.reset_args()


def _run(calls, fuse):
    aq.set_args(variant="A", bitW=8, abitW=8, act_range=2, fuse_bn_act=True, own_conv="tf32", own_conv_channels=(16,),
                method="none", fuse_dgrad_bn=fuse)
    C = 16
    blocks = torch.nn.ModuleList([PreActBlock_conv_Q("second", 8, 8, C, C, 1, variant="A") for _ in range(2)]).train()
    for m in blocks.modules():
        if hasattr(m, "quantize_fn"):
            m.quantize_fn = torch.nn.Identity()          # the weight quantizer needs the GPU; not what is tested here
    bn0 = torch.nn.BatchNorm2d(C).train()
    q0 = aq.activation_quantize_fn(8, "second")
    x = torch.randn(2, C, 4, 4).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    out = fused.bn_act(bn0, q0, x, True)
    for blk in blocks:
        out = blk(out)
    del calls[:]
    out.sum().backward()
    return list(calls)


def test_fused_backward_call_sequence(recorder):
    seq = _run(recorder, True)
    assert seq == ["alignq_bn_act_bwd_sum"] + ["alignq_conv3x3_bwd_weight", "alignq_conv3x3_bwd_data_bnreduce",
                                               "alignq_bn_act_bwd_apply"] * 4, seq


def test_unfused_backward_call_sequence(recorder):
    seq = _run(recorder, False)
    assert seq == ["alignq_bn_act_bwd_sum"] + ["alignq_conv3x3_bwd_weight", "alignq_conv3x3_bwd_data",
                                               "alignq_bn_act_bwd_sum"] * 4, seq
