"""CPU/torch-eager restatement of the reference's CIFAR ResNet (QA and QB forms) and of one
training iteration of its ``main.py`` -- built ONLY on ``oracle/alignq_oracle.py`` and stock torch.

TEST INFRASTRUCTURE ONLY (see the header of alignq_oracle.py).  Used by the tests as the model-level
checker and by ``bench.py --impl reference`` / ``cpu_baseline`` as the timed CPU baseline
(kind "port": the reference itself is Python and cannot travel to the GPU box).

Follows  cdf_alignment/resnet-20-cifar-10/model/resnet.py:33-138            (variant A)
         cdf_alignment_admm/resnet-56-cifar-10/model/resnet.py:36-167        (variant B)
         cdf_alignment/resnet-20-cifar-10/main.py:269-313, cdf_alignment_admm/resnet-56-cifar-10/main.py:286-379
Pinned by oracle/make_model_golden.py: same state_dict -> bit-identical logits / trans_loss / grads
as the imported reference on CPU.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import alignq_oracle as O


class OracleADMM(nn.Module):                                     # utils/admm.py:12-33
    def __init__(self, dim):
        super().__init__()
        self.mu, self.rho = 0.2, 0.3
        self.alterD = nn.Parameter(torch.rand(dim, dim))
        self.gamma = nn.Parameter(torch.rand(dim, dim))
        self.D = None


class OracleConv(nn.Conv2d):                                     # quantization.py:107-122
    def __init__(self, cin, cout, k, stride, padding, w_bit, variant, groups=1):
        super().__init__(cin, cout, k, stride, padding, groups=groups, bias=False)
        self.w_bit, self.variant = w_bit, variant
        self.weight_cdf = self.weight_pdf = None

    def forward(self, x):
        wq, self.weight_cdf, self.weight_pdf = O.weight_quantize(self.weight, self.w_bit, self.variant)
        return F.conv2d(x, wq, None, self.stride, self.padding, self.dilation, self.groups)


class OracleAct(nn.Module):
    def __init__(self, a_bit, variant, act_range, admm=None):
        super().__init__()
        self.a_bit, self.variant, self.act_range = a_bit, variant, act_range
        self.opt = admm

    def forward(self, x):
        if self.opt is None:
            return O.activation_quantize(x, self.a_bit, "second", self.variant, self.act_range)
        y, loss, D = O.activation_quantize_admm(x, self.a_bit, self.opt.alterD, self.opt.gamma, "second",
                                                self.variant, self.act_range, self.opt.mu, self.opt.rho)
        self.opt.D = D
        return y, loss


class OracleBlock(nn.Module):
    def __init__(self, wbit, abit, cin, cout, stride, variant, ar, dim):
        super().__init__()
        admm = variant != "A"
        if admm:
            self.admm0, self.admm1 = OracleADMM(dim), OracleADMM(dim)
        self.act_q0 = OracleAct(abit, variant, ar, self.admm0 if admm else None)
        self.act_q1 = OracleAct(abit, variant, ar, self.admm1 if admm else None)
        self.bn0 = nn.BatchNorm2d(cout)
        self.conv0 = OracleConv(cin, cout, 3, stride, 1, wbit, variant)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv1 = OracleConv(cout, cout, 3, 1, 1, wbit, variant)
        self.skip_conv = None
        if stride != 1:
            if admm:
                self.admm_skip = OracleADMM(dim)
            self.act_skip_q = OracleAct(abit, variant, ar, self.admm_skip if admm else None)
            self.skip_conv = OracleConv(cin, cout, 1, stride, 0, wbit, variant)
            self.skip_bn = nn.BatchNorm2d(cout)
        self.admm = admm

    def forward(self, x):
        tl = 0.
        shortcut = x
        if self.skip_conv is not None:
            shortcut = self.act_skip_q(self.skip_bn(self.skip_conv(x)))
            if self.admm:
                shortcut, l = shortcut
                tl += l
        out = self.act_q0(self.bn0(self.conv0(x)))
        if self.admm:
            out, l = out
            tl += l
        out = F.relu(out)
        out = self.act_q1(self.bn1(self.conv1(out)))
        if self.admm:
            out, l = out
            tl += l
        out += shortcut
        out = F.relu(out)
        return (out, tl) if self.admm else out


class OracleResNet(nn.Module):
    def __init__(self, units, wbit, abit, variant="A", act_range=2.0, dim=128, num_classes=10):
        super().__init__()
        self.admm = variant != "A"
        self.conv0 = OracleConv(3, 16, 3, 1, 1, wbit, variant)
        if self.admm:
            self.admm0 = OracleADMM(dim)
        self.act_q0 = OracleAct(abit, variant, act_range, self.admm0 if self.admm else None)
        self.layers = nn.ModuleList()
        cin = 16
        for stage, width in enumerate((16, 32, 64)):
            for u in range(units[stage]):
                stride = 2 if (stage > 0 and u == 0) else 1
                self.layers.append(OracleBlock(wbit, abit, cin, width, stride, variant, act_range, dim))
                cin = width
        self.bn = nn.BatchNorm2d(16)
        self.logit = nn.Linear(64, num_classes)

    def forward(self, x):
        tl = 0.
        out = self.act_q0(self.bn(self.conv0(x)))
        if self.admm:
            out, l = out
            tl += l
        out = F.relu(out)
        for layer in self.layers:
            out = layer(out)
            if self.admm:
                out, l = out
                tl += l
        out = F.adaptive_avg_pool2d(out, 1).view(out.size(0), -1)
        out = self.logit(out)
        return (out, tl) if self.admm else out


def resnet20_oracle(wbit, abit, variant="A", **kw):
    return OracleResNet([3, 3, 3], wbit, abit, variant, **kw)


def resnet56_oracle(wbit, abit, variant="A", **kw):
    return OracleResNet([9, 9, 9], wbit, abit, variant, **kw)


# ---- DenseNet-40 (variant A): cdf_alignment/dense-cifar-10/model/densenet.py:17-159 -------------------------------------
class OracleDenseBlock(nn.Module):                               # DenseBasicBlock, densenet.py:17-42
    def __init__(self, wbit, abit, inplanes, growth, ar):
        super().__init__()
        self.act_q0 = OracleAct(abit, "A", ar)
        self.bn1 = nn.BatchNorm2d(inplanes)
        self.conv1 = OracleConv(inplanes, growth, 3, 1, 1, wbit, "A")

    def forward(self, x):
        out = self.conv1(F.relu(self.act_q0(self.bn1(x))))
        return torch.cat((x, out), 1)


class OracleTransition(nn.Module):                               # Transition, densenet.py:45-63
    def __init__(self, wbit, abit, inplanes, outplanes, ar):
        super().__init__()
        self.act_q0 = OracleAct(abit, "A", ar)
        self.bn1 = nn.BatchNorm2d(inplanes)
        self.conv1 = OracleConv(inplanes, outplanes, 1, 1, 0, wbit, "A")

    def forward(self, x):
        return F.avg_pool2d(self.conv1(F.relu(self.act_q0(self.bn1(x)))), 2)


class OracleDenseNet(nn.Module):                                 # DenseNet(depth=40, compressionRate=1), densenet.py:66-159
    admm = False

    def __init__(self, wbit, abit, act_range=2.0, depth=40, growth=12, num_classes=10):
        super().__init__()
        n = (depth - 4) // 3
        planes = 2 * growth
        self.conv1 = OracleConv(3, planes, 3, 1, 1, wbit, "A")
        stages = []
        for st in range(3):
            blocks = []
            for _ in range(n):
                blocks.append(OracleDenseBlock(wbit, abit, planes, growth, act_range))
                planes += growth
            stages.append(nn.Sequential(*blocks))
            if st < 2:
                stages.append(OracleTransition(wbit, abit, planes, planes, act_range))      # compressionRate = 1
        self.dense1, self.trans1, self.dense2, self.trans2, self.dense3 = stages
        self.bn = nn.BatchNorm2d(planes)
        self.fc = nn.Linear(planes, num_classes)
        self.act_q0 = OracleAct(abit, "A", act_range)

    def forward(self, x):
        x = self.dense3(self.trans2(self.dense2(self.trans1(self.dense1(self.conv1(x))))))
        x = F.relu(self.act_q0(self.bn(x)))
        return self.fc(F.avg_pool2d(x, 8).view(x.size(0), -1))

    def quant_convs(self):
        """The convolutions in the order dense-cifar-10/main.py:303-316 collects weight_cdf / weight_pdf."""
        convs = [self.conv1]
        for j, layer in enumerate((self.dense1, self.trans1, self.dense2, self.trans2, self.dense3)):
            convs += [blk.conv1 for blk in layer] if j % 2 == 0 else [layer.conv1]
        return convs

    skip_first_conv = False                                      # main.py:297-300: every conv weight takes the surrogate


# ---- MobileNet-v2 (variant A): cdf_alignment/mobilenet-v2-svhn/model/mobilenetV2.py:23-130 ---------------------------
class OracleMBBlock(nn.Module):                                  # Block, mobilenetV2.py:23-71
    def __init__(self, wbit, abit, cin, cout, expansion, stride, ar):
        super().__init__()
        self.stride = stride
        planes = expansion * cin
        self.act_q1, self.act_q2, self.act_q3 = (OracleAct(abit, "A", ar) for _ in range(3))
        self.act_skip = OracleAct(abit, "A", ar)
        self.conv1 = OracleConv(cin, planes, 1, 1, 0, wbit, "A")
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = OracleConv(planes, planes, 3, stride, 1, wbit, "A", groups=planes)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = OracleConv(planes, cout, 1, 1, 0, wbit, "A")
        self.bn3 = nn.BatchNorm2d(cout)
        self.shortcut = None
        if stride == 1:
            self.shortcut = nn.Sequential(OracleConv(cin, cout, 1, 1, 0, wbit, "A"), nn.BatchNorm2d(cout), self.act_skip, nn.ReLU())

    def forward(self, x):
        out = F.relu6(self.act_q1(self.bn1(self.conv1(x))))
        out = F.relu6(self.act_q2(self.bn2(self.conv2(out))))
        out = self.act_q3(self.bn3(self.conv3(out)))
        if self.stride == 1:
            out += self.shortcut(x)
        return out


class OracleMobileNetV2(nn.Module):                              # MobileNetV2, mobilenetV2.py:74-130
    admm = False
    cfg = [(1, 16, 1, 1), (6, 24, 2, 1), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]

    def __init__(self, wbit, abit, act_range=2.0, num_classes=10):
        super().__init__()
        self.act_q1, self.act_q2 = OracleAct(abit, "A", act_range), OracleAct(abit, "A", act_range)
        self.conv1 = OracleConv(3, 32, 3, 1, 1, wbit, "A")
        self.bn1 = nn.BatchNorm2d(32)
        layers, cin = [], 32
        for expansion, cout, nblocks, stride in self.cfg:
            for st in [stride] + [1] * (nblocks - 1):
                layers.append(OracleMBBlock(wbit, abit, cin, cout, expansion, st, act_range))
                cin = cout
        self.layers = nn.Sequential(*layers)
        self.conv2 = OracleConv(320, 1280, 1, 1, 0, wbit, "A")
        self.bn2 = nn.BatchNorm2d(1280)
        self.linear = nn.Linear(1280, num_classes)

    def forward(self, x):
        out = F.relu(self.act_q1(self.bn1(self.conv1(x))))
        out = self.layers(out)
        out = F.relu(self.act_q2(self.bn2(self.conv2(out))))
        out = F.avg_pool2d(out, 4)
        return self.linear(out.view(out.size(0), -1))

    def quant_convs(self):
        """The convolutions in the order mobilenet-v2-svhn/main.py:182-199 collects weight_cdf / weight_pdf."""
        convs = [self.conv1]
        for layer in self.layers:
            convs += [layer.conv1, layer.conv2, layer.conv3] + ([layer.shortcut[0]] if layer.shortcut is not None else [])
        return convs + [self.conv2]

    skip_first_conv = False                                      # main.py:176-180 (`idx = idx[1:]` is commented out there)

    @staticmethod
    def quant_weight_name(n):                                    # main.py:177
        return ("conv" in n and "weight" in n) or ("shortcut.0" in n and "weight" in n)


def deterministic_fill(state_dict, seed=0):
    """Return a copy of the state_dict filled with seeded values that depend only on key order and shape, so the
    reference, the oracle and the product can be given identical weights without shipping them."""
    # clone first: the reference registers every ADMM module twice (admm0.* and act_q0.opt.*, same
    # storage); with private copies the winner is decided by load_state_dict's traversal order, which
    # is identical for the reference, the oracle and the product (same registration order)
    state_dict = {k: v.detach().clone() for k, v in state_dict.items()}
    g = torch.Generator().manual_seed(seed)
    for k in sorted(state_dict.keys()):
        t = state_dict[k]
        if not t.is_floating_point():
            t.zero_()
        elif k.endswith(("alterD", "gamma")):
            t.copy_(torch.rand(t.shape, generator=g))
        elif "running_var" in k:
            t.fill_(1.0)
        elif "running_mean" in k:
            t.zero_()
        elif t.dim() == 1 and "bias" not in k:                  # BN weight
            t.copy_(1.0 + 0.1 * torch.randn(t.shape, generator=g))
        elif t.dim() == 1:
            t.copy_(0.05 * torch.randn(t.shape, generator=g))
        else:
            fan = t[0].numel()
            t.copy_(torch.randn(t.shape, generator=g) * (2.0 / fan) ** 0.5)
    return state_dict


class OracleTrainer:
    """One iteration of the reference's train() body on synthetic tensors (main.py:269-313 for QA,
    cdf_alignment_admm/.../main.py:286-379 for QB), with the optimizers restated functionally."""

    def __init__(self, model, lr=0.04, momentum=0.9, weight_decay=1e-4, lam=1.0, lam2=4.0, bitW=8):
        self.model = model
        self.named = [(n, p) for n, p in model.named_parameters() if "alterD" not in n and "gamma" not in n]
        self.bufs = [None] * len(self.named)
        self.hp = dict(lr=lr, momentum=momentum, weight_decay=weight_decay)
        self.lam, self.lam2, self.bitW = lam, lam2, bitW
        is_q = getattr(model, "quant_weight_name", lambda n: "conv" in n and "weight" in n)
        idx = [j for j, (n, _) in enumerate(self.named) if is_q(n)]
        if hasattr(model, "quant_convs"):                        # DenseNet: its own collection order, every conv
            self.idx = idx[1:] if model.skip_first_conv else idx
            self.convs = model.quant_convs()
        else:
            self.idx = idx[1:]                                   # main.py:299-304
            self.convs = [c for layer in model.layers for c in (layer.conv0, layer.conv1, layer.skip_conv) if c is not None]
        self.admms = []
        if model.admm:
            self.admms.append(model.admm0)
            for layer in model.layers:
                self.admms += [layer.admm0, layer.admm1] + ([layer.admm_skip] if layer.skip_conv is not None else [])

    def step(self, inputs, targets):
        m = self.model
        for p in m.parameters():
            p.grad = None
        out = m(inputs)
        if m.admm:
            out, tl = out
            ce = F.cross_entropy(out, targets)
            ce.backward(retain_graph=True)                       # .../main.py:300
            tl = tl + 0.5
            tl.backward()
        else:
            ce = F.cross_entropy(out, targets)
            ce.backward()
        with torch.no_grad():
            params = [p for _, p in self.named]
            grads = [p.grad for p in params]
            w_cdf = [c.weight_cdf.detach() for c in self.convs]
            w_pdf = [c.weight_pdf.detach() for c in self.convs]
            new_g = O.sgd_step(params, grads, self.bufs, self.idx, w_cdf, w_pdf, self.lam, self.lam2,
                               bitW=self.bitW, **self.hp)
            for p, g in zip(params, new_g):
                p.grad = g
            for a in self.admms:
                if a.alterD.grad is None:
                    continue
                Z, U = O.admm_zu_update(a.D.detach(), a.alterD.detach(), a.gamma.detach(), a.mu, a.rho)
                a.alterD.copy_(Z)
                a.gamma.copy_(U)
        return ce.detach(), out.detach()
