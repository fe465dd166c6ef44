"""The thin data side the YAML front door needs — NOT a rebuild of torchok/data.

The reference's data pipeline (albumentations transforms, csv / COCO / SOP datasets, samplers) is outside the hot
path and outside this package's scope.  What is here is what lets `examples/configs/classification_cifar10.yaml`
load and run unchanged on an offline box:

* TRANSFORMS `Compose`, `Resize`, `Normalize`, `ToTensorV2`, `HorizontalFlip`, `VerticalFlip`, `CenterCrop`,
  `RandomCrop` with albumentations' calling convention (`t(image=..., mask=...) -> dict`) and defaults
  (`Normalize`: `(img / 255 - mean) / std`; `Resize`: cv2 bilinear for images, nearest for masks);
* DATASETS `CIFAR10` / `CIFAR100` reading the standard python pickles from `data_folder` (sample dict of
  torchok/data/datasets/examples/cifar.py:139-163: 'image' CHW tensor of `input_dtype`, 'target', 'index').  There is
  no network here: `download: true` with the files absent raises the reference's RuntimeError text plus a hint;
* `SyntheticImages`: seeded random images / targets of a given shape, for smoke runs and benchmarks;
* `create_dataloaders(data_cfg, phase)` = Constructor.create_dataloaders (torchok/constructor/constructor.py:268-312).
"""
import os
import pickle

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

from ..constructor import DATASETS, SAMPLERS, TRANSFORMS

try:  # cv2 is only needed by Resize
    import cv2
except ImportError:  # pragma: no cover
    cv2 = None


# ---------------------------------------------------------------------------------------------------- transforms
class _Transform:
    def __init__(self, always_apply=False, p=1.0):
        self.p = 1.0 if always_apply else p

    def image(self, img, **params):
        return img

    def mask(self, m, **params):
        return self.image(m, **params)

    def params(self, sample):
        return {}

    def __call__(self, **sample):
        if self.p < 1.0 and np.random.random() >= self.p:
            return sample
        prm = self.params(sample)
        out = dict(sample)
        if 'image' in out:
            out['image'] = self.image(out['image'], **prm)
        if out.get('mask') is not None:
            out['mask'] = self.mask(out['mask'], **prm)
        return out


@TRANSFORMS.register_class
class Compose:
    def __init__(self, transforms, p=1.0, **unused):
        self.transforms, self.p = list(transforms), p

    def __call__(self, **sample):
        for t in self.transforms:
            sample = t(**sample)
        return sample


@TRANSFORMS.register_class
class OneOf:
    """albumentations.OneOf: with probability `p` apply exactly one of `transforms`, drawn with weights equal to the
    members' own `p` (normalised)."""

    def __init__(self, transforms, p=0.5, **unused):
        self.transforms, self.p = list(transforms), p
        w = np.array([getattr(t, 'p', 1.0) for t in self.transforms], dtype=np.float64)
        self.weights = w / w.sum() if w.sum() > 0 else np.full(len(w), 1.0 / max(len(w), 1))

    def __call__(self, **sample):
        if self.transforms and np.random.random() < self.p:
            t = self.transforms[int(np.random.choice(len(self.transforms), p=self.weights))]
            saved, t.p = getattr(t, 'p', 1.0), 1.0       # the chosen member is applied unconditionally
            try:
                sample = t(**sample)
            finally:
                t.p = saved
        return sample


@TRANSFORMS.register_class
class Resize(_Transform):
    def __init__(self, height, width, interpolation=1, always_apply=False, p=1.0):
        super().__init__(always_apply, p)
        self.height, self.width, self.interpolation = height, width, interpolation

    def _resize(self, img, interpolation):
        if img.shape[:2] == (self.height, self.width):
            return img
        if cv2 is None:
            raise ImportError('Resize needs opencv (cv2)')
        return cv2.resize(img, (self.width, self.height), interpolation=interpolation)

    def image(self, img, **params):
        return self._resize(img, self.interpolation)

    def mask(self, m, **params):
        return self._resize(m, 0)


@TRANSFORMS.register_class
class Normalize(_Transform):
    def __init__(self, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), max_pixel_value=255.0,
                 always_apply=False, p=1.0):
        super().__init__(always_apply, p)
        self.mean = np.asarray(mean, dtype=np.float32) * max_pixel_value
        self.inv = np.reciprocal(np.asarray(std, dtype=np.float32) * max_pixel_value)

    def image(self, img, **params):
        return (img.astype(np.float32) - self.mean) * self.inv

    def mask(self, m, **params):
        return m


@TRANSFORMS.register_class
class ToTensorV2(_Transform):
    def __init__(self, transpose_mask=False, always_apply=True, p=1.0):
        super().__init__(True, 1.0)
        self.transpose_mask = transpose_mask

    def image(self, img, **params):
        if img.ndim == 2:
            img = img[:, :, None]
        return torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1)))

    def mask(self, m, **params):
        if self.transpose_mask and m.ndim == 3:
            m = m.transpose(2, 0, 1)
        return torch.from_numpy(np.ascontiguousarray(m))


@TRANSFORMS.register_class
class HorizontalFlip(_Transform):
    def __init__(self, always_apply=False, p=0.5):
        super().__init__(always_apply, p)

    def image(self, img, **params):
        return np.ascontiguousarray(img[:, ::-1])


@TRANSFORMS.register_class
class VerticalFlip(_Transform):
    def __init__(self, always_apply=False, p=0.5):
        super().__init__(always_apply, p)

    def image(self, img, **params):
        return np.ascontiguousarray(img[::-1])


@TRANSFORMS.register_class
class CenterCrop(_Transform):
    def __init__(self, height, width, always_apply=False, p=1.0):
        super().__init__(always_apply, p)
        self.height, self.width = height, width

    def params(self, sample):
        h, w = sample['image'].shape[:2]
        if h < self.height or w < self.width:
            raise ValueError(f'Requested crop size ({self.height}, {self.width}) is larger than the image size ({h}, {w})')
        return {'y0': (h - self.height) // 2, 'x0': (w - self.width) // 2}

    def image(self, img, y0=0, x0=0):
        return img[y0:y0 + self.height, x0:x0 + self.width]


@TRANSFORMS.register_class
class RandomCrop(CenterCrop):
    def params(self, sample):
        h, w = sample['image'].shape[:2]
        if h < self.height or w < self.width:
            raise ValueError(f'Requested crop size ({self.height}, {self.width}) is larger than the image size ({h}, {w})')
        return {'y0': int(np.random.random() * (h - self.height + 1)), 'x0': int(np.random.random() * (w - self.width + 1))}


def create_transforms(specs):
    """Constructor._create_transforms (constructor.py:332-365): a list of {name, params} → Compose; a transform whose
    params hold `transforms` is a container built recursively."""
    if specs is None:
        return None

    def build(items):
        out = []
        for item in items:
            params = dict(item.get('params') or {})
            if 'transforms' in params:
                inner = build(params.pop('transforms'))
                out.append(TRANSFORMS.get(item['name'])(transforms=inner, **params))
            else:
                out.append(TRANSFORMS.get(item['name'])(**params))
        return out
    return TRANSFORMS.get('Compose')(transforms=build(specs))


# ------------------------------------------------------------------------------------------------------ datasets
class ImageDataset(Dataset):
    """The constructor surface of torchok/data/datasets/base.py:16-50 (transform / augment / input_dtype / test_mode)."""

    def __init__(self, transform, augment=None, input_dtype='float32', reader_library='opencv', image_format='rgb',
                 rgba_layout_color=0, test_mode=False):
        self.transform, self.augment, self.input_dtype, self.test_mode = transform, augment, input_dtype, test_mode
        self.reader_library, self.image_format, self.rgba_layout_color = reader_library, image_format, rgba_layout_color

    @staticmethod
    def _apply(transform, sample):
        return sample if transform is None else transform(**sample)

    def get_raw(self, idx):
        raise NotImplementedError

    def __getitem__(self, idx):
        sample = self._apply(self.transform, self.get_raw(idx))
        sample['image'] = sample['image'].type(getattr(torch, self.input_dtype))
        return sample


@DATASETS.register_class
class CIFAR10(ImageDataset):
    base_folder = 'cifar-10-batches-py'
    train_list = ['data_batch_1', 'data_batch_2', 'data_batch_3', 'data_batch_4', 'data_batch_5']
    test_list = ['test_batch']
    meta = ('batches.meta', 'label_names')

    def __init__(self, train, download, data_folder, transform, augment=None, input_dtype='float32',
                 reader_library='opencv', image_format='rgb', rgba_layout_color=0, test_mode=False):
        super().__init__(transform, augment, input_dtype, reader_library, image_format, rgba_layout_color, test_mode)
        self.train = train
        root = os.path.join(str(data_folder), self.base_folder)
        names = self.train_list if train else self.test_list
        missing = [n for n in self.train_list + self.test_list + [self.meta[0]]
                   if not os.path.exists(os.path.join(root, n))]
        if missing:
            raise RuntimeError('Dataset not found or corrupted. You can use download=True to download it'
                               f' [torchok_b200: no network on this box, download is not attempted; expected '
                               f'{", ".join(missing)} under {root}]')
        images, targets = [], []
        for name in names:
            with open(os.path.join(root, name), 'rb') as f:
                entry = pickle.load(f, encoding='latin1')
            images.append(entry['data'])
            targets.extend(entry['labels'] if 'labels' in entry else entry['fine_labels'])
        self.targets = np.array(targets, dtype=np.int64)
        self.images = np.vstack(images).reshape(-1, 3, 32, 32).transpose(0, 2, 3, 1)   # HWC
        with open(os.path.join(root, self.meta[0]), 'rb') as f:
            self.classes = pickle.load(f, encoding='latin1')[self.meta[1]]
        self.class_to_idx = {c: i for i, c in enumerate(self.classes)}

    def get_raw(self, idx):
        sample = {'image': self.images[idx], 'index': idx}
        if not self.test_mode:
            sample['target'] = self.targets[idx]
        return self._apply(self.augment, sample)

    def __len__(self):
        return len(self.images)


@DATASETS.register_class
class CIFAR100(CIFAR10):
    base_folder = 'cifar-100-python'
    train_list = ['train']
    test_list = ['test']
    meta = ('meta', 'fine_label_names')


@DATASETS.register_class
class SyntheticImages(ImageDataset):
    """Seeded random uint8 images with classification (`target`: int64 scalar) or segmentation (`target`: H×W int64,
    `task='segmentation'`) labels.  Not in the reference; it exists because the box has no datasets and no network."""

    def __init__(self, transform=None, augment=None, num_samples=1024, shape=(32, 32, 3), num_classes=10, seed=0,
                 task='classification', input_dtype='float32', test_mode=False, **unused):
        super().__init__(transform, augment, input_dtype, test_mode=test_mode)
        rng = np.random.RandomState(seed)
        h, w, c = shape
        self.images = rng.randint(0, 256, size=(num_samples, h, w, c), dtype=np.uint8)
        self.task = task
        if task == 'segmentation':
            self.targets = rng.randint(0, num_classes, size=(num_samples, h, w)).astype(np.int64)
        else:
            self.targets = rng.randint(0, num_classes, size=(num_samples,)).astype(np.int64)

    def get_raw(self, idx):
        sample = {'image': self.images[idx], 'index': idx}
        if not self.test_mode:
            sample['mask' if self.task == 'segmentation' else 'target'] = self.targets[idx]
        return self._apply(self.augment, sample)

    def __getitem__(self, idx):
        if self.transform is None:
            sample = self.get_raw(idx)
            sample['image'] = torch.from_numpy(sample['image'].transpose(2, 0, 1).astype(np.float32) / 255.0)
            if 'mask' in sample:
                sample['mask'] = torch.from_numpy(sample['mask'])
        else:
            sample = self._apply(self.transform, self.get_raw(idx))
        sample['image'] = sample['image'].type(getattr(torch, self.input_dtype))
        if 'mask' in sample:
            sample['target'] = sample.pop('mask').long()
        return sample

    def __len__(self):
        return len(self.images)


# ---------------------------------------------------------------------------------------------------- dataloaders
def create_dataset(dataset_cfg):
    transform = create_transforms(dataset_cfg.get('transform'))
    augment = create_transforms(dataset_cfg.get('augment'))
    return DATASETS.get(dataset_cfg['name'])(transform=transform, augment=augment, **dict(dataset_cfg.get('params') or {}))


def create_dataloaders(data_cfg, phase, distributed_sampler=True):
    """List of DataLoaders for `phase` ('TRAIN' | 'VALID' | 'TEST' | 'PREDICT'); [] when the phase has no entry.
    Under torch.distributed a DistributedSampler shards the dataset (Lightning's `use_distributed_sampler`)."""
    import torch.distributed as dist
    if not data_cfg or phase not in data_cfg or not data_cfg[phase]:
        return []
    loaders = []
    for entry in data_cfg[phase]:
        if entry is None:
            continue
        dataset = create_dataset(entry['dataset'])
        params = dict(entry.get('dataloader') or {})
        sampler = None
        if entry.get('sampler') is not None:
            sp = dict(entry['sampler'].get('params') or {})
            sp.setdefault('num_samples', len(dataset))
            sampler = SAMPLERS.get(entry['sampler']['name'])(**sp)
        elif distributed_sampler and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            from torch.utils.data.distributed import DistributedSampler
            sampler = DistributedSampler(dataset, shuffle=bool(params.pop('shuffle', False)),
                                         drop_last=bool(params.get('drop_last', False)))
        if sampler is not None:
            params.pop('shuffle', None)
        if params.get('num_workers', 0) == 0:
            params.pop('prefetch_factor', None)
            params.pop('persistent_workers', None)
        params.setdefault('pin_memory', True)
        loaders.append(DataLoader(dataset, collate_fn=getattr(dataset, 'collate_fn', None), sampler=sampler, **params))
    return loaders
