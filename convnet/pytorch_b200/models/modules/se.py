"""Squeeze-and-excitation gate of the reference (models/modules/se.py:6-25): global average pool -> Linear(C, C/ratio)
-> ReLU -> Linear(C/ratio, C) -> Sigmoid, multiplied onto the input.  In resnet_se / resnext_se it sits on the RESIDUAL
branch of every block (models/resnet.py:112-113,159-160) and ONE instance is shared by all blocks of a stage
(models/resnet.py:182-191).  Attribute names match the reference so checkpoints interchange."""
import torch.nn as nn


class SEBlock(nn.Module):
    def __init__(self, in_channels, out_channels=None, ratio=16):
        super(SEBlock, self).__init__()
        self.in_channels = in_channels
        out_channels = in_channels if out_channels is None else out_channels
        self.ratio = ratio
        self.relu = nn.ReLU(True)
        self.global_pool = nn.AdaptiveAvgPool2d(1)
        self.transform = nn.Sequential(nn.Linear(in_channels, in_channels // ratio), nn.ReLU(inplace=True),
                                       nn.Linear(in_channels // ratio, out_channels), nn.Sigmoid())

    def forward(self, x):
        gate = self.transform(self.global_pool(x).flatten(1, -1))
        return x * gate.unsqueeze(-1).unsqueeze(-1)
