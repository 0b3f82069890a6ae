"""MobileNet-v2 on the reference's factory contract (models/mobilenet_v2.py:39-165).

Inverted residual = 1x1 expand (x6) -> depthwise 3x3 -> 1x1 linear projection, BN after each, ReLU6 after
the first two, identity skip when stride 1 and the widths match.  Module names (``features.conv0``,
``features.bottleneck{i}.block.{j}``, ``features.conv1``, ``classifier.1``) match the reference so
checkpoints interchange.  Weight decay skips depthwise convolutions (models/mobilenet_v2.py:25-36).
"""
import math

import torch.nn as nn

__all__ = ['mobilenet_v2']


def nearby_int(n):
    return int(round(n))


def init_model(model):
    for m in model.modules():
        if isinstance(m, nn.Conv2d):
            fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
            m.weight.data.normal_(0, math.sqrt(2. / fan_out))
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.data.fill_(1)
            m.bias.data.zero_()


def weight_decay_config(value=1e-4, log=False):
    def decayed(m):
        dense_conv = isinstance(m, nn.Conv2d) and m.groups != m.in_channels
        return dense_conv or isinstance(m, nn.Linear)
    return {'name': 'WeightDecay', 'value': value, 'log': log,
            'filter': {'parameter_name': lambda n: not n.endswith('bias'), 'module': decayed}}


def conv_bn_relu6(cin, cout, kernel=3, stride=1, padding=1, groups=1):
    return [nn.Conv2d(cin, cout, kernel, stride, padding, groups=groups, bias=False), nn.BatchNorm2d(cout),
            nn.ReLU6(inplace=True)]


class ExpandedConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, expansion=1, kernel_size=3, stride=1, padding=1,
                 residual_block=None):
        super(ExpandedConv2d, self).__init__()
        hidden = in_channels * expansion
        self.add_res = stride == 1 and in_channels == out_channels
        self.residual_block = residual_block
        layers = [] if hidden == in_channels else conv_bn_relu6(in_channels, hidden, 1, 1, 0)
        layers += conv_bn_relu6(hidden, hidden, kernel_size, stride, padding, groups=hidden)
        layers += [nn.Conv2d(hidden, out_channels, 1, bias=False), nn.BatchNorm2d(out_channels)]
        self.block = nn.Sequential(*layers)

    def forward(self, x):
        out = self.block(x)
        if self.add_res:
            out = out + (x if self.residual_block is None else self.residual_block(x))
        return out


# (expansion, stride, base width) for the 17 inverted-residual blocks
_BLOCKS = [(1, 1, 16), (6, 2, 24), (6, 1, 24), (6, 2, 32), (6, 1, 32), (6, 1, 32), (6, 2, 64), (6, 1, 64),
           (6, 1, 64), (6, 1, 64), (6, 1, 96), (6, 1, 96), (6, 1, 96), (6, 2, 160), (6, 1, 160), (6, 1, 160),
           (6, 1, 320)]


class MobileNet_v2(nn.Module):
    _b200 = None

    def __init__(self, width=1., regime=None, num_classes=1000, scale_lr=1):
        super(MobileNet_v2, self).__init__()
        cin = nearby_int(width * 32)
        self.features = nn.Sequential()
        self.features.add_module('conv0', nn.Sequential(*conv_bn_relu6(3, cin, 3, 2, 1)))
        for i, (t, s, c) in enumerate(_BLOCKS):
            cout = nearby_int(width * c)
            self.features.add_module('bottleneck%d' % i, ExpandedConv2d(cin, cout, expansion=t, stride=s))
            cin = cout
        last = nearby_int(width * 1280)
        self.features.add_module('conv1', nn.Sequential(*conv_bn_relu6(cin, last, 1, 1, 0)))
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.classifier = nn.Sequential(nn.Dropout(0.2, True), nn.Linear(last, num_classes))
        init_model(self)

        if regime == 'small':
            scale_lr *= 4
            self.data_regime = [{'epoch': 0, 'input_size': 128, 'batch_size': 512},
                                {'epoch': 80, 'input_size': 224, 'batch_size': 128}]
            self.data_eval_regime = [{'epoch': 0, 'input_size': 128, 'scale_size': 160, 'batch_size': 1024},
                                     {'epoch': 80, 'input_size': 224, 'batch_size': 512}]
        self.regime = [{'epoch': 0, 'optimizer': 'SGD', 'momentum': 0.9, 'lr': scale_lr * 1e-1,
                        'regularizer': weight_decay_config(1e-4)}] + \
                      [{'epoch': e, 'lr': scale_lr * 10 ** -(k + 2)} for k, e in enumerate((30, 60, 80))]

    def forward(self, x):
        if self._b200 is not None:
            return self._b200.forward(x)
        x = self.avg_pool(self.features(x)).flatten(1)
        return self.classifier(x)


def mobilenet_v2(**config):
    dataset = config.pop('dataset', 'imagenet')
    use_b200 = config.pop('b200', False)
    assert 'imagenet' in dataset, 'mobilenet_v2 is defined for ImageNet-shaped inputs'
    model = MobileNet_v2(**config)
    if use_b200:
        from ..engine import convert_b200
        model = convert_b200(model)
    return model
