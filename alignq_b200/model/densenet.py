"""DenseNet-40 (k = 12) hosting the quantized modules (config 4 of BASELINE.json).
Topology and names follow cdf_alignment/dense-cifar-10/model/densenet.py:17-159."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .fused import bn_act
from .quantization import activation_quantize_fn, conv2d_Q_fn


class DenseBasicBlock(nn.Module):
    def __init__(self, stage, wbit, abit, inplanes, filters, growthRate=12, dropRate=0, variant=None):
        super().__init__()
        Conv2d = conv2d_Q_fn(w_bit=wbit, stage=stage, variant=variant)
        self.act_q0 = activation_quantize_fn(abit, stage, variant=variant)
        self.bn1 = nn.BatchNorm2d(inplanes)
        self.conv1 = Conv2d(filters, growthRate, kernel_size=3, padding=1, bias=False)
        self.relu = nn.ReLU(inplace=True)
        self.dropRate = dropRate

    def forward(self, x):
        out = self.conv1(bn_act(self.bn1, self.act_q0, x, True))
        if self.dropRate > 0:
            out = F.dropout(out, p=self.dropRate, training=self.training)
        return torch.cat((x, out), 1)


class Transition(nn.Module):
    def __init__(self, stage, wbit, abit, inplanes, outplanes, filters, variant=None):
        super().__init__()
        Conv2d = conv2d_Q_fn(w_bit=wbit, stage=stage, variant=variant)
        self.act_q0 = activation_quantize_fn(abit, stage, variant=variant)
        self.bn1 = nn.BatchNorm2d(inplanes)
        self.conv1 = Conv2d(filters, outplanes, kernel_size=1, bias=False)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        return F.avg_pool2d(self.conv1(bn_act(self.bn1, self.act_q0, x, True)), 2)


class DenseNet(nn.Module):
    def __init__(self, wbit, abit, stage, depth=40, dropRate=0, num_classes=10, growthRate=12, compressionRate=2,
                 variant=None):
        super().__init__()
        assert (depth - 4) % 3 == 0, "depth should be 3n+4"
        n = (depth - 4) // 3
        self.wbit, self.abit, self.stage, self.variant = wbit, abit, stage, variant
        self.growthRate, self.dropRate = growthRate, dropRate
        Conv2d = conv2d_Q_fn(w_bit=wbit, stage=stage, variant=variant)
        self.inplanes = growthRate * 2
        self.conv1 = Conv2d(3, self.inplanes, kernel_size=3, padding=1, bias=False)
        self.dense1 = self._dense(n)
        self.trans1 = self._transition(compressionRate)
        self.dense2 = self._dense(n)
        self.trans2 = self._transition(compressionRate)
        self.dense3 = self._dense(n)
        self.bn = nn.BatchNorm2d(self.inplanes)
        self.relu = nn.ReLU(inplace=True)
        self.avgpool = nn.AvgPool2d(8)
        self.fc = nn.Linear(self.inplanes, num_classes)
        self.act_q0 = activation_quantize_fn(abit, stage, variant=variant)
        for m in self.modules():                                  # densenet.py:111-118
            if isinstance(m, nn.Conv2d):
                fan = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / fan))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _dense(self, blocks):
        layers = []
        for _ in range(blocks):
            layers.append(DenseBasicBlock(self.stage, self.wbit, self.abit, self.inplanes, self.inplanes,
                                          self.growthRate, self.dropRate, self.variant))
            self.inplanes += self.growthRate
        return nn.Sequential(*layers)

    def _transition(self, compressionRate):
        inplanes = self.inplanes
        self.inplanes = int(math.floor(self.inplanes // compressionRate))
        return Transition(self.stage, self.wbit, self.abit, inplanes, self.inplanes, inplanes, self.variant)

    def forward(self, x):
        x = self.dense3(self.trans2(self.dense2(self.trans1(self.dense1(self.conv1(x))))))
        x = self.avgpool(bn_act(self.bn, self.act_q0, x, True))
        return self.fc(x.view(x.size(0), -1))


def densenet_40_quant(bitW, abitW, stage, pretrained=False, **kwargs):
    return DenseNet(wbit=bitW, abit=abitW, stage=stage, depth=40, compressionRate=1, **kwargs)
