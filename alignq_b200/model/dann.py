"""ResNet-50 DANN for Office-31 hosting the quantized modules (config 5 of BASELINE.json).
Follows cdf_alignment_admm/dann_office/model/resnet.py:29-334 (Bottleneck / ResNet / ReverseLayerF / DANN);
only the resnet50 path is provided (the reference's BasicBlock variants are broken as shipped,
SURVEY.md A.5 #7).  No pretrained download: weights are random-init or loaded by the caller."""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.autograd import Function

from ..utils.admm import ADMM
from ..utils.options import args
from .quantization import activation_quantize_fn, activation_quantize_fn2, conv2d_Q_fn


def conv3x3(wbit, stage, in_planes, out_planes, stride=1, groups=1, dilation=1, variant="C"):
    return conv2d_Q_fn(wbit, stage, variant)(in_planes, out_planes, kernel_size=3, stride=stride, padding=dilation,
                                              groups=groups, bias=False, dilation=dilation)


def conv1x1(wbit, stage, in_planes, out_planes, stride=1, variant="C"):
    return conv2d_Q_fn(wbit, stage, variant)(in_planes, out_planes, kernel_size=1, stride=stride, bias=False)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, wbit, abit, stage, inplanes, planes, stride=1, downsample=None, variant="C"):
        super().__init__()
        width = planes
        self.conv1 = conv1x1(wbit, stage, inplanes, width, variant=variant)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = conv3x3(wbit, stage, width, width, stride, variant=variant)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = conv1x1(wbit, stage, width, planes * self.expansion, variant=variant)
        self.bn3 = nn.BatchNorm2d(planes * self.expansion)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride
        self.admm0 = ADMM(args.train_batch_size)
        self.act_q1 = activation_quantize_fn(abit, stage, variant=variant)
        self.act_q2 = activation_quantize_fn(abit, stage, variant=variant)
        self.act_q3 = activation_quantize_fn2(abit, stage, self.admm0, variant=variant)

    def forward(self, x):
        identity = x
        out = self.relu(self.act_q1(self.bn1(self.conv1(x))))
        out = self.relu(self.act_q2(self.bn2(self.conv2(out))))
        out, trans_loss = self.act_q3(self.bn3(self.conv3(out)))
        if self.downsample is not None:
            identity = self.downsample(x)
        out += identity
        return self.relu(out), 0. + trans_loss


class ResNet(nn.Module):
    def __init__(self, wbit, abit, stage, layers, num_classes=1000, variant="C"):
        super().__init__()
        self.wbit, self.abit, self.stage, self.variant = wbit, abit, stage, variant
        self.act_q0 = activation_quantize_fn(abit, stage, variant=variant)
        self.inplanes = 64
        self.conv1 = conv2d_Q_fn(wbit, stage, variant)(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(64, layers[0])
        self.layer2 = self._make_layer(128, layers[1], stride=2)
        self.layer3 = self._make_layer(256, layers[2], stride=2)
        self.layer4 = self._make_layer(512, layers[3], stride=2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512 * Bottleneck.expansion, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _make_layer(self, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * Bottleneck.expansion:
            downsample = nn.Sequential(
                conv1x1(self.wbit, self.stage, self.inplanes, planes * Bottleneck.expansion, stride, self.variant),
                nn.BatchNorm2d(planes * Bottleneck.expansion))
        layers = [Bottleneck(self.wbit, self.abit, self.stage, self.inplanes, planes, stride, downsample, self.variant)]
        self.inplanes = planes * Bottleneck.expansion
        for _ in range(1, blocks):
            layers.append(Bottleneck(self.wbit, self.abit, self.stage, self.inplanes, planes, variant=self.variant))
        return nn.Sequential(*layers)

    def forward(self, x):
        trans_loss = 0.
        x = self.maxpool(self.relu(self.act_q0(self.bn1(self.conv1(x)))))
        for stage_layers in (self.layer1, self.layer2, self.layer3, self.layer4):
            for layer in stage_layers:
                x, loss = layer(x)
                trans_loss += loss
        feature = torch.flatten(self.avgpool(x), 1)
        return feature, trans_loss


def resnet50_quant(wbit, abit, stage, pretrained=False, progress=True, **kwargs):
    if pretrained:
        raise RuntimeError("alignq_b200 ships no pretrained download; load a state_dict yourself")
    return ResNet(wbit, abit, stage, [3, 4, 6, 3], **kwargs)


class ReverseLayerF(Function):
    @staticmethod
    def forward(ctx, x, alpha):
        ctx.alpha = alpha
        return x.view_as(x)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.neg() * ctx.alpha, None


class DANN(nn.Module):
    def __init__(self, arch, wbit, abit, stage, num_classes=31):
        super().__init__()
        self.feature = arch(wbit, abit, stage)
        self.class_classifier = nn.Sequential()
        self.class_classifier.add_module("c_fc3", nn.Linear(2048, num_classes))
        self.domain_classifier = nn.Sequential()
        self.domain_classifier.add_module("d_fc2", nn.Linear(2048, 2))

    def forward(self, input_data, alpha):
        feature, trans_loss = self.feature(input_data)
        feature = feature.view(-1, 2048)
        reverse_feature = ReverseLayerF.apply(feature, alpha)
        return self.class_classifier(feature), self.domain_classifier(reverse_feature), trans_loss


def resnet50_dann(wbit, abit, stage, **kwargs):
    return DANN(resnet50_quant, wbit, abit, stage)
