"""ResNeXt = the ResNet constructors with grouped 3x3 convolutions (reference: models/resnext.py:10-55).

``resnext(depth=101)`` with the defaults below is ResNeXt-101 32x4d (width 128..1024, 32 groups,
expansion 2).
"""
from .resnet import ResNet_imagenet, ResNet_cifar, BasicBlock, Bottleneck

__all__ = ['resnext', 'resnext_se']

_LAYERS = {18: (BasicBlock, [2, 2, 2, 2]), 34: (BasicBlock, [3, 4, 6, 3]), 50: (Bottleneck, [3, 4, 6, 3]),
           101: (Bottleneck, [3, 4, 23, 3]), 152: (Bottleneck, [3, 8, 36, 3])}


class ResNeXt_imagenet(ResNet_imagenet):
    def __init__(self, width=[128, 256, 512, 1024], groups=[32, 32, 32, 32], expansion=2, **kwargs):
        super(ResNeXt_imagenet, self).__init__(width=width, groups=groups, expansion=expansion, **kwargs)


class ResNeXt_cifar(ResNet_cifar):
    def __init__(self, width=[64, 128, 256], groups=[4, 8, 16], **kwargs):
        super(ResNeXt_cifar, self).__init__(width=width, groups=groups, **kwargs)


def resnext(**config):
    dataset = config.pop('dataset', 'imagenet')
    use_b200 = config.pop('b200', False)
    if 'imagenet' in dataset:
        config.setdefault('num_classes', 1000)
        depth = config.pop('depth', 50)
        if depth in _LAYERS:
            block, layers = _LAYERS[depth]
            config.update(block=block, layers=layers)
        model = ResNeXt_imagenet(**config)
    elif 'cifar' in dataset:
        config.setdefault('num_classes', 100 if '100' in dataset else 10)
        config.setdefault('depth', 44)
        model = ResNeXt_cifar(block=BasicBlock, **config)
    else:
        raise ValueError('resnext: unknown dataset %r' % dataset)
    if use_b200:
        from ..engine import convert_b200
        model = convert_b200(model)
    return model


def resnext_se(**config):
    """ResNeXt with squeeze-and-excitation gates (models/resnext.py:53-55 of the reference)."""
    from .modules.se import SEBlock
    config['residual_block'] = SEBlock
    return resnext(**config)
