"""CIFAR pre-activation ResNets hosting the quantized modules (configs 1 and 2 of BASELINE.json).

Topology, attribute names and registration order follow the reference so that state_dict keys and
``named_parameters()`` order (which the reference's training loop indexes into) are identical:
  variant 'A'      cdf_alignment/resnet-20-cifar-10/model/resnet.py:33-138   forward -> logits
  variant 'B'/'C'  cdf_alignment_admm/resnet-56-cifar-10/model/resnet.py:36-167 forward -> (logits, trans_loss)
"""
from __future__ import annotations

import torch.nn as nn
import torch.nn.functional as F

from ..utils.admm import ADMM
from ..utils.options import args
from .fused import avgpool_linear_ce, bn_act, conv_bn_act, fork, side_branch
from .quantization import activation_quantize_fn, conv2d_Q_fn


def _with_admm(variant):
    return (args.variant if variant is None else variant) != "A"


class PreActBlock_conv_Q(nn.Module):
    """conv-bn-actq-relu-conv-bn-actq (+ quantized 1x1 projection shortcut when stride != 1)."""

    def __init__(self, stage, wbit, abit, in_planes, out_planes, stride=1, variant=None):
        super().__init__()
        Conv2d = conv2d_Q_fn(w_bit=wbit, stage=stage, variant=variant)
        self.with_admm = _with_admm(variant)
        if self.with_admm:
            dim = args.train_batch_size                       # resnet.py:43-46 (always the train branch)
            self.admm0, self.admm1 = ADMM(dim), ADMM(dim)
            self.act_q0 = activation_quantize_fn(abit, stage, self.admm0, variant=variant)
            self.act_q1 = activation_quantize_fn(abit, stage, self.admm1, variant=variant)
        else:
            self.act_q0 = activation_quantize_fn(abit, stage, variant=variant)
            self.act_q1 = activation_quantize_fn(abit, stage, variant=variant)
            self.act_skip_q = activation_quantize_fn(abit, stage, variant=variant)
        self.bn0 = nn.BatchNorm2d(out_planes)
        self.conv0 = Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(out_planes)
        self.conv1 = Conv2d(out_planes, out_planes, kernel_size=3, stride=1, padding=1, bias=False)
        self.skip_conv = None
        if stride != 1:
            if self.with_admm:
                self.admm_skip = ADMM(args.train_batch_size)
                self.act_skip_q = activation_quantize_fn(abit, stage, self.admm_skip, variant=variant)
            self.skip_conv = Conv2d(in_planes, out_planes, kernel_size=1, stride=stride, padding=0, bias=False)
            self.skip_bn = nn.BatchNorm2d(out_planes)

    def forward(self, x):
        if not self.with_admm:
            x, xs = fork(x)           # two consumers: their gradients meet inside the producer's backward kernel
            if self.skip_conv is None:
                shortcut = xs
                out = conv_bn_act(self.conv0, self.bn0, self.act_q0, x, True)  # relu(act_q0(bn0(conv0(x))))
            else:
                # the projection shortcut (1x1 conv -> bn -> act-quant) does not depend on conv0 / bn0: inside a training
                # step it runs on a side stream beside them (autograd runs its backward on that stream as well)
                with side_branch(xs) as br:
                    shortcut = bn_act(self.skip_bn, self.act_skip_q, self.skip_conv(xs), False)
                out = conv_bn_act(self.conv0, self.bn0, self.act_q0, x, True)
                br.join(shortcut)
            return conv_bn_act(self.conv1, self.bn1, self.act_q1, out, True, residual=shortcut)   # relu(act_q1(.) + shortcut)
        trans_loss = 0.
        shortcut = x
        if self.skip_conv is not None:
            shortcut, loss = self.act_skip_q(self.skip_bn(self.skip_conv(x)))
            trans_loss += loss
        out, loss = self.act_q0(self.bn0(self.conv0(x)))
        trans_loss += loss
        out, loss = self.act_q1(self.bn1(self.conv1(F.relu(out))))
        trans_loss += loss
        out += shortcut
        return F.relu(out), trans_loss


class PreActResNet(nn.Module):
    def __init__(self, block, num_units, wbit, abit, stage, num_classes, block_bits=None, variant=None):
        super().__init__()
        Conv2d = conv2d_Q_fn(w_bit=wbit, stage=stage, variant=variant)
        self.with_admm = _with_admm(variant)
        self.conv0 = Conv2d(3, 16, kernel_size=3, stride=1, padding=1, bias=False)
        if self.with_admm:
            self.admm0 = ADMM(args.train_batch_size)
            self.act_q0 = activation_quantize_fn(abit, stage, self.admm0, variant=variant)
        else:
            self.act_q0 = activation_quantize_fn(abit, stage, variant=variant)
        self.layers = nn.ModuleList()
        widths = [16] * num_units[0] + [32] * num_units[1] + [64] * num_units[2]
        strides = [1] * num_units[0] + [2] + [1] * (num_units[1] - 1) + [2] + [1] * (num_units[2] - 1)
        in_planes = 16
        for n, (stride, width) in enumerate(zip(strides, widths)):
            bits = wbit if block_bits is None else block_bits[n]
            self.layers.append(block(stage, bits, abit, in_planes, width, stride, variant=variant))
            in_planes = width
        self.bn = nn.BatchNorm2d(16)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.logit = nn.Linear(64, num_classes)

    def features(self, x):
        """Everything up to (not including) the average pool: (feature map, trans_loss or None)."""
        if not self.with_admm:
            out = conv_bn_act(self.conv0, self.bn, self.act_q0, x, True)
            for layer in self.layers:
                out = layer(out)
            return out, None
        trans_loss = 0.
        out, loss = self.act_q0(self.bn(self.conv0(x)))
        trans_loss += loss
        out = F.relu(out)
        for layer in self.layers:
            out, loss = layer(out)
            trans_loss += loss
        return out, trans_loss

    def forward(self, x):
        out, trans_loss = self.features(x)
        logits = self.logit(self.avgpool(out).view(out.size(0), -1))
        return logits if trans_loss is None else (logits, trans_loss)

    def forward_ce(self, x, target):
        """forward + ``F.cross_entropy(output, target)`` with the pool -> linear -> loss tail on the fused head kernels
        (model/fused.py:avgpool_linear_ce): (loss, detached logits, trans_loss or None)."""
        out, trans_loss = self.features(x)
        loss, logits = avgpool_linear_ce(out, self.logit, target)
        return loss, logits, trans_loss


def resnet20_quant(bitW, abitW, stage, num_classes=10, variant=None):
    return PreActResNet(PreActBlock_conv_Q, [3, 3, 3], bitW, abitW, stage, num_classes=num_classes, variant=variant)


def resnet56_quant(bitW, abitW, stage, num_classes=10, variant=None):
    return PreActResNet(PreActBlock_conv_Q, [9, 9, 9], bitW, abitW, stage, num_classes=num_classes, variant=variant)
