"""MobileNet-v2 (CIFAR/SVHN 32x32 form) hosting the quantized modules incl. depthwise convs
(config 3 of BASELINE.json).  Follows cdf_alignment/mobilenet-v2-svhn/model/mobilenetV2.py:25-137."""
from __future__ import annotations

import torch.nn as nn

from ..utils.options import args
from .fused import bn_act
from .quantization import activation_quantize_fn, conv2d_Q_fn


class Block(nn.Module):
    """expand 1x1 -> depthwise 3x3 -> project 1x1, each followed by BN + activation quantizer."""

    def __init__(self, stage, wbit, abit, in_planes, out_planes, expansion, stride, variant=None):
        super().__init__()
        self.stride = stride
        planes = expansion * in_planes
        Conv2d = conv2d_Q_fn(w_bit=wbit, stage=stage, variant=variant)
        self.act_q1 = activation_quantize_fn(abit, stage, variant=variant)
        self.act_q2 = activation_quantize_fn(abit, stage, variant=variant)
        self.act_q3 = activation_quantize_fn(abit, stage, variant=variant)
        self.act_skip = activation_quantize_fn(abit, stage, variant=variant)
        self.conv1 = Conv2d(in_planes, planes, kernel_size=1, stride=1, padding=0, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = Conv2d(planes, planes, kernel_size=3, stride=stride, padding=1, groups=planes, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = Conv2d(planes, out_planes, kernel_size=1, stride=1, padding=0, bias=False)
        self.bn3 = nn.BatchNorm2d(out_planes)
        self.relu = nn.ReLU6()
        self.shortcut = None
        if stride == 1:
            self.shortcut = nn.Sequential(
                Conv2d(in_planes, out_planes, kernel_size=1, stride=1, padding=0, bias=False),
                nn.BatchNorm2d(out_planes), self.act_skip, nn.ReLU(inplace=True))

    def forward(self, x):
        if args.act_range <= 6 and self.act_q1.a_bit < 32:   # |act_q| <= act_range: ReLU6 == ReLU on the quantizer's output
            out = bn_act(self.bn1, self.act_q1, self.conv1(x), True)
            out = bn_act(self.bn2, self.act_q2, self.conv2(out), True)
        else:
            out = self.relu(self.act_q1(self.bn1(self.conv1(x))))
            out = self.relu(self.act_q2(self.bn2(self.conv2(out))))
        out = bn_act(self.bn3, self.act_q3, self.conv3(out), False)
        if self.stride == 1:
            out += self.shortcut(x)
        return out


class MobileNetV2(nn.Module):
    cfg = [(1, 16, 1, 1), (6, 24, 2, 1), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]

    def __init__(self, block, wbit, abit, stage, num_classes=10, variant=None):
        super().__init__()
        Conv2d = conv2d_Q_fn(w_bit=wbit, stage=stage, variant=variant)
        self.act_q1 = activation_quantize_fn(abit, stage, variant=variant)
        self.act_q2 = activation_quantize_fn(abit, stage, variant=variant)
        self.conv1 = Conv2d(3, 32, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(32)
        layers, in_planes = [], 32
        for expansion, out_planes, num_blocks, stride in self.cfg:
            for s in [stride] + [1] * (num_blocks - 1):
                layers.append(block(stage, wbit, abit, in_planes, out_planes, expansion, s, variant=variant))
                in_planes = out_planes
        self.layers = nn.Sequential(*layers)
        self.conv2 = Conv2d(320, 1280, kernel_size=1, stride=1, padding=0, bias=False)
        self.bn2 = nn.BatchNorm2d(1280)
        self.linear = nn.Linear(1280, num_classes)
        self.relu = nn.ReLU(inplace=True)
        self.avg_pool2d = nn.AvgPool2d(4)

    def forward(self, x):
        out = bn_act(self.bn1, self.act_q1, self.conv1(x), True)
        out = self.layers(out)
        out = bn_act(self.bn2, self.act_q2, self.conv2(out), True)
        self.out = self.avg_pool2d(out)
        return self.linear(self.out.view(self.out.size(0), -1))


def mobile_v2(wbit, abit, stage, **kwargs):
    return MobileNetV2(Block, wbit, abit, stage, **kwargs)
