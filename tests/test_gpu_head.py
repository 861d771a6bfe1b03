"""GPU: the fused classifier head + loss (csrc/head_ce.cu; model/resnet.py:108-110 + main.py:283-286 of the reference's
resnet-20-cifar-10): loss / logits against torch in fp32, gradients against the fp64 autograd of the same chain at 1e-5."""
import pytest
import torch
import torch.nn.functional as F

import alignq_b200 as aq
from alignq_b200.model.fused import avgpool_linear_ce, head_ce_applies

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _reset():
    yield
    aq.reset_args()


def relmax(a, b):
    return float((a.detach().double() - b.detach().double()).abs().max() / b.detach().double().abs().max())


@pytest.mark.parametrize("B,C,H,K,bias", [(128, 64, 8, 10, True), (5, 64, 7, 31, True), (3, 256, 1, 1000, False), (32, 12, 4, 3, True),
                                          (1, 1024, 2, 10, True), (200, 64, 8, 10, True)])
def test_head_ce_matches_pool_linear_cross_entropy(B, C, H, K, bias):
    torch.manual_seed(70)
    feat0 = torch.randn(B, C, H, H, device=DEV).contiguous(memory_format=torch.channels_last)
    lin = torch.nn.Linear(C, K, bias=bias).to(DEV)
    t = torch.randint(0, K, (B,), device=DEV)
    assert head_ce_applies(feat0, lin, t)
    for rep in range(2):                                            # second call: re-armed ticket
        lin.zero_grad()
        feat = feat0.clone().requires_grad_(True)
        loss, logits = avgpool_linear_ce(feat, lin, t)
        (loss * 1.7).backward()
    f64 = feat0.double().clone().requires_grad_(True)
    w64 = lin.weight.detach().double().clone().requires_grad_(True)
    b64 = lin.bias.detach().double().clone().requires_grad_(True) if bias else None
    lg64 = F.linear(F.avg_pool2d(f64, H).view(B, -1), w64, b64)
    l64 = F.cross_entropy(lg64, t)
    (l64 * 1.7).backward()
    assert not logits.requires_grad and loss.shape == ()
    assert abs(float(loss.detach()) - float(l64.detach())) <= 2e-6 * abs(float(l64.detach())) + 1e-7
    assert relmax(logits, lg64) <= 2e-6
    assert feat.grad.is_contiguous(memory_format=torch.channels_last) or H == 1
    assert relmax(feat.grad, f64.grad) <= 1e-5
    assert relmax(lin.weight.grad, w64.grad) <= 1e-5
    if bias:
        assert relmax(lin.bias.grad, b64.grad) <= 1e-5


def test_head_ce_falls_back_to_the_library_chain_when_it_does_not_apply():
    feat = torch.randn(4, 6, 3, 3, device=DEV, requires_grad=True)          # C % 4 != 0
    lin = torch.nn.Linear(6, 5).to(DEV)
    t = torch.randint(0, 5, (4,), device=DEV)
    assert not head_ce_applies(feat, lin, t)
    loss, logits = avgpool_linear_ce(feat, lin, t)
    ref = F.cross_entropy(lin(F.avg_pool2d(feat, 3).view(4, -1)), t)
    assert torch.allclose(loss, ref)


def test_qat_step_with_the_fused_head_follows_the_unfused_step():
    from alignq_b200.model import resnet
    from alignq_b200.utils.train import QATStep
    B = 32
    torch.manual_seed(71)
    x = torch.randn(B, 3, 32, 32, device=DEV).contiguous(memory_format=torch.channels_last)
    t = torch.randint(0, 10, (B,), device=DEV)
    res = {}
    for fused in (False, True):
        aq.reset_args()
        aq.set_args(variant="A", train_batch_size=B, bitW=8, abitW=8, act_range=2, fused_head=fused)
        torch.manual_seed(0)
        m = resnet.resnet20_quant(8, 8, "second").to(DEV).train()
        st = QATStep(m, lr=0.04, momentum=0.9, weight_decay=1e-4, channels_last=True, keep_logits=True)
        w0, b0 = m.logit.weight.detach().clone(), m.logit.bias.detach().clone()
        loss = float(st.step(x, t))
        res[fused] = (loss, m.logit.weight.detach() - w0, m.logit.bias.detach() - b0, st.logits.clone())
    assert abs(res[True][0] - res[False][0]) <= 2e-6 * abs(res[False][0])
    assert relmax(res[True][3], res[False][3]) <= 1e-5
    # the classifier's own update depends on the forward pass only (everything upstream of it goes through the quantizers'
    # straight-through gradients, where a 1e-7 difference flips rounding ties: see test_graph_replay_equals_eager)
    assert relmax(res[True][1], res[False][1]) <= 1e-4 and relmax(res[True][2], res[False][2]) <= 1e-4
