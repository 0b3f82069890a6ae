"""Data regimes: the (input size, batch size, duplicates) schedule side of the reference's data.py:74-203.

``DataRegime(regime, defaults)`` resolves, per epoch, the dataset / transform / loader settings from a regime
list (same grammar as the optimizer regimes) and hands out a ``DataLoader``; ``SampledDataRegime`` mixes
several regimes with given probabilities (Mix&Match, models/resnet.py:279-311 in the reference).

Datasets: the north-star configurations are synthetic, so ``synthetic_*`` names are first class here
(the reference has no synthetic option and needs downloads -- SURVEY.md section 8c shim 5):
  synthetic_cifar10 / synthetic_cifar100 : N(0,1) 3x32x32 tensors, 10 / 100 classes
  synthetic_imagenet                     : N(0,1) 3xSxS tensors (S = input_size, default 224), 1000 classes
Real datasets (cifar10, cifar100, imagenet folders) go through torchvision if it is importable; the PIL
augmentation stack of the reference (preprocess.py, autoaugment.py) is out of scope.
"""
import os
from copy import deepcopy
from itertools import accumulate, chain

import torch
from torch.utils.data import Dataset, Subset
from torch.utils.data.distributed import DistributedSampler

from .utils.regime import Regime


class SyntheticImages(Dataset):
    """Deterministic N(0,1) images with uniform random labels; sample i is a pure function of (seed, i).

    A small pool of ``pool`` distinct images is generated once and indexed modulo, so that iterating costs
    nothing next to a training step (the metric excludes data loading: SURVEY.md section 8d)."""

    def __init__(self, length, shape, num_classes, seed=0, duplicates=1, pool=256):
        g = torch.Generator().manual_seed(seed)
        self.length = length
        self.duplicates = duplicates
        self.pool = min(pool, length)
        self.images = torch.randn(self.pool, *shape, generator=g)
        self.labels = torch.randint(0, num_classes, (length,), generator=g)

    def __len__(self):
        return self.length

    def __getitem__(self, idx):
        img = self.images[idx % self.pool]
        if self.duplicates > 1:  # batch augmentation: D "augmentations" of the sample (here: D pool neighbours)
            img = torch.stack([self.images[(idx + d) % self.pool] for d in range(self.duplicates)])
        return img, int(self.labels[idx])


_IMAGE_STATS = {'mean': [0.485, 0.456, 0.406], 'std': [0.229, 0.224, 0.225]}   # preprocess.py:7-8 of the reference


def real_dataset_transform(transform_name='imagenet', input_size=None, scale_size=None, normalize=None, augment=True,
                           cutout=None, autoaugment=False, padding=None, duplicates=1, num_crops=1):
    """The basic transform stacks the reference's ``get_transform`` selects for real image datasets
    (preprocess.py:114-161): random-resized-crop + flip (imagenet, train), pad-4 random crop + flip (cifar, train),
    resize + centre crop (evaluation), then ToTensor + Normalize.  The research augmentations -- AutoAugment, Cutout,
    Lighting/colour jitter, multi-crop evaluation -- are outside this repo's scope
    and raise instead of being silently dropped."""
    import torchvision.transforms as T
    if autoaugment or cutout is not None or num_crops != 1:
        raise NotImplementedError('autoaugment / cutout / multi-crop evaluation are outside the B200 hot path '
                                  '(reference preprocess.py, autoaugment.py); use the reference data pipeline')
    stats = normalize or _IMAGE_STATS
    tail = [T.ToTensor(), T.Normalize(**stats)]
    if 'imagenet' in transform_name:
        input_size = input_size or 224
        scale_size = scale_size or int(input_size * 8 / 7)
        if augment:
            # the reference's inception_preprocess adds ColorJitter + PCA Lighting (preprocess.py:77-97) which are not
            # reproduced: say so rather than train on a silently different distribution
            import logging
            logging.warning('imagenet training transform: RandomResizedCrop + flip only (no colour jitter / lighting)')
            head = [T.RandomResizedCrop(input_size), T.RandomHorizontalFlip()]
        else:
            head = ([T.Resize(scale_size)] if scale_size != input_size else []) + [T.CenterCrop(input_size)]
    elif 'cifar' in transform_name:
        input_size = input_size or 32
        scale_size = scale_size or 32
        if augment:
            head = [T.RandomCrop(scale_size, padding=padding or 4)]
            if input_size != scale_size:
                head.append(T.Resize(input_size))
            head.append(T.RandomHorizontalFlip())
        else:
            head = ([T.Resize(scale_size)] if scale_size != input_size else []) + [T.CenterCrop(input_size)]
    else:
        raise NotImplementedError('no transform for dataset family %r' % transform_name)
    fn = T.Compose(head + tail)
    if duplicates > 1:   # batch augmentation: D independent draws of the transform per sample (preprocess.py:105-112)
        return T.Lambda(lambda img: torch.stack([fn(img) for _ in range(duplicates)], dim=0))
    return fn


_SYNTHETIC = {'synthetic_cifar10': (32, 10, 50000, 10000), 'synthetic_cifar100': (32, 100, 50000, 10000),
              'synthetic_imagenet': (224, 1000, 1281167, 50000)}


def get_dataset(name, split='train', transform=None, target_transform=None, download=True,
                datasets_path='~/Datasets', input_size=None, duplicates=1, synthetic_length=None):
    train = split == 'train'
    if name in _SYNTHETIC:
        size, classes, n_train, n_val = _SYNTHETIC[name]
        size = input_size or size
        length = synthetic_length or int(os.environ.get('B200_SYNTHETIC_LENGTH', n_train if train else n_val))
        return SyntheticImages(length, (3, size, size), classes, seed=0 if train else 1, duplicates=duplicates)
    try:
        import torchvision.datasets as tvd
        import torchvision.transforms as T
    except ImportError as e:  # pragma: no cover
        raise ValueError('dataset %r needs torchvision (%s); use a synthetic_* dataset' % (name, e))
    root = os.path.join(os.path.expanduser(datasets_path), name)
    if transform is None:
        transform = T.ToTensor()
    if name == 'cifar10':
        return tvd.CIFAR10(root=root, train=train, transform=transform, target_transform=target_transform,
                           download=download)
    if name == 'cifar100':
        return tvd.CIFAR100(root=root, train=train, transform=transform, target_transform=target_transform,
                            download=download)
    if name == 'imagenet':
        return tvd.ImageFolder(root=os.path.join(root, 'train' if train else 'val'), transform=transform,
                               target_transform=target_transform)
    raise ValueError('unknown dataset %r' % name)


_DATA_ARGS = {'name', 'split', 'transform', 'target_transform', 'download', 'datasets_path', 'synthetic_length'}
_DATALOADER_ARGS = {'batch_size', 'shuffle', 'sampler', 'batch_sampler', 'num_workers', 'collate_fn', 'pin_memory',
                    'drop_last', 'timeout', 'worker_init_fn'}
_TRANSFORM_ARGS = {'transform_name', 'input_size', 'scale_size', 'normalize', 'augment', 'cutout', 'duplicates',
                   'num_crops', 'autoaugment'}
_OTHER_ARGS = {'distributed'}


class DataRegime(object):
    def __init__(self, regime, defaults={}):
        self.regime = Regime(regime, deepcopy(defaults))
        self.epoch = 0
        self.steps = None
        self._sampler = None
        self.get_loader(True)

    def get_setting(self):
        setting = self.regime.setting
        pick = lambda keys: {k: v for k, v in setting.items() if k in keys}  # noqa: E731
        out = {'data': pick(_DATA_ARGS), 'loader': pick(_DATALOADER_ARGS), 'transform': pick(_TRANSFORM_ARGS),
               'other': pick(_OTHER_ARGS)}
        out['transform'].setdefault('transform_name', out['data'].get('name'))
        return out

    def get(self, key, default=None):
        return self.regime.setting.get(key, default)

    def get_loader(self, force_update=False, override_settings=None, subset_indices=None):
        if force_update or self.regime.update(self.epoch, self.steps):
            setting = self.get_setting()
            if override_settings is not None:
                setting.update(override_settings)
            data_kwargs = dict(setting['data'])
            name = data_kwargs.get('name', '')
            if name in _SYNTHETIC:  # the "transform" of a synthetic dataset is just its geometry
                data_kwargs['input_size'] = setting['transform'].get('input_size')
                data_kwargs['duplicates'] = setting['transform'].get('duplicates') or 1
            elif data_kwargs.get('transform') is None:
                # real images: build the transform the regime asks for (reference data.py:101-102) -- never fall back
                # to a bare ToTensor(), which would train on unnormalised, unaugmented, variable-size images
                tf = {k: v for k, v in setting['transform'].items() if v is not None}
                data_kwargs['transform'] = real_dataset_transform(**tf)
            self._data = get_dataset(**data_kwargs)
            if subset_indices is not None:
                self._data = Subset(self._data, subset_indices)
            loader_kwargs = dict(setting['loader'])
            if name in _SYNTHETIC:
                loader_kwargs['num_workers'] = 0
            if setting['other'].get('distributed', False):
                loader_kwargs['sampler'] = DistributedSampler(self._data)
                loader_kwargs['shuffle'] = None
                loader_kwargs['pin_memory'] = False
            self._sampler = loader_kwargs.get('sampler', None)
            self._loader = torch.utils.data.DataLoader(self._data, **loader_kwargs)
        return self._loader

    def set_epoch(self, epoch):
        self.epoch = epoch
        if self._sampler is not None and hasattr(self._sampler, 'set_epoch'):
            self._sampler.set_epoch(epoch)

    def __len__(self):
        return len(self._data)

    def __repr__(self):
        return str(self.regime)


class SampledDataLoader(object):
    """Interleaves several loaders; the order is a permutation seeded by the epoch only, so that all ranks
    draw the same (size, batch) sequence (data.py:133-139 in the reference)."""

    def __init__(self, dl_list):
        self.dl_list = dl_list
        self.epoch = 0

    def generate_order(self):
        order = list(chain(*[[idx] * len(dl) for idx, dl in enumerate(self.dl_list)]))
        g = torch.Generator().manual_seed(self.epoch)
        return torch.tensor(order)[torch.randperm(len(order), generator=g)].tolist()

    def __len__(self):
        return sum(len(dl) for dl in self.dl_list)

    def __iter__(self):
        iterators = [iter(dl) for dl in self.dl_list]
        for idx in self.generate_order():
            yield next(iterators[idx])


class SampledDataRegime(DataRegime):
    def __init__(self, data_regime_list, probs, split_data=True):
        self.probs = probs
        self.data_regime_list = data_regime_list
        self.split_data = split_data
        self.epoch = 0

    def get_setting(self):
        return [r.get_setting() for r in self.data_regime_list]

    def get(self, key, default=None):
        return [r.get(key, default) for r in self.data_regime_list]

    def get_loader(self, force_update=False):
        if self.split_data:
            sizes = {len(r._data.dataset) if isinstance(r._data, Subset) else len(r._data)
                     for r in self.data_regime_list}
            assert len(sizes) == 1, 'all datasets should be same size'
            total = sizes.pop()
            lengths = [int(p * total) for p in self.probs]
            lengths[-1] = total - sum(lengths[:-1])
            g = torch.Generator().manual_seed(1000003 + self.epoch)  # identical split on every rank
            perm = torch.randperm(total, generator=g).tolist()
            parts = [perm[end - n:end] for end, n in zip(accumulate(lengths), lengths)]
            loaders = [r.get_loader(force_update=True, subset_indices=parts[i])
                       for i, r in enumerate(self.data_regime_list)]
        else:
            loaders = [r.get_loader(force_update=force_update) for r in self.data_regime_list]
        self._loader = SampledDataLoader(loaders)
        self._loader.epoch = self.epoch
        return self._loader

    def set_epoch(self, epoch):
        self.epoch = epoch
        if hasattr(self, '_loader'):
            self._loader.epoch = epoch
        for r in self.data_regime_list:
            if r._sampler is not None and hasattr(r._sampler, 'set_epoch'):
                r._sampler.set_epoch(epoch)

    def __len__(self):
        return sum(len(r._data) for r in self.data_regime_list)

    def __repr__(self):
        return 'Sampled Data Regime:\n' + ''.join('w.p. %s:  %s\n' % (p, r)
                                                   for p, r in zip(self.probs, self.data_regime_list))
