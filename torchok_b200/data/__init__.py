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
* DATASETS `ImageClassificationDataset`, `ImageSegmentationDataset`, `SOP`, `SweetPepper` (bottom of this file) with
  the reference's annotation formats and sample dicts; TRANSFORMS `OneOf`, `FancyPCA`;
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


# ------------------------------------------------------------------------------------- file-backed example datasets
# What the other hot-path example configs name (classification_imagenet.yaml, pairwise_sop.yaml,
# segmentation_sweet_pepper.yaml).  Same constructor arguments, annotation formats and sample dicts as the reference;
# archives are never downloaded (no network): a missing folder raises the reference's RuntimeError text.
def _read_image(path, image_format='rgb', reader_library='opencv', rgba_layout_color=0):
    """torchok/data/datasets/base.py:67-92 for the formats the example datasets hold (8-bit gray / RGB / RGBA files)."""
    if cv2 is None:
        raise ImportError('reading image files needs opencv (cv2)')
    img = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
    if img is None:
        raise ValueError(f'{path} image does not exist')
    if img.dtype != np.uint8:
        img = (img // 256).astype('uint8')
    if img.ndim == 3 and img.shape[2] == 4:
        img = cv2.cvtColor(img, cv2.COLOR_BGRA2RGBA)
        alpha = img[..., 3:4] / 255
        img = np.clip(img[..., :3] * alpha + rgba_layout_color * (1 - alpha), 0, 255).astype('uint8')
    elif img.ndim == 3:
        img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
    if image_format == 'rgb':
        return cv2.cvtColor(img, cv2.COLOR_GRAY2RGB) if img.ndim == 2 else img
    if image_format == 'bgr':
        return cv2.cvtColor(img, cv2.COLOR_GRAY2BGR) if img.ndim == 2 else cv2.cvtColor(img, cv2.COLOR_RGB2BGR)
    if image_format == 'gray':
        return (img if img.ndim == 2 else cv2.cvtColor(img, cv2.COLOR_RGB2GRAY))[..., None]
    raise ValueError(f'Unsupported image format `{image_format}`')


def _read_table(folder, annotation_path, dtype=None, **kw):
    import pandas as pd
    path = os.path.join(str(folder), annotation_path)
    if annotation_path.endswith('.csv') or annotation_path.endswith('.txt'):
        return pd.read_csv(path, dtype=dtype, **kw)
    if annotation_path.endswith('.pkl'):
        return pd.read_pickle(path)
    raise ValueError('Detection dataset error. Annotation path is not in `csv` or `pkl` format')


@DATASETS.register_class
class ImageClassificationDataset(ImageDataset):
    """csv / pkl of (image_path, label) rows: torchok/data/datasets/classification/classification.py:38-209."""

    def __init__(self, data_folder, transform, augment=None, annotation_path=None, num_classes=None,
                 input_column='image_path', input_dtype='float32', target_column='label', target_dtype='long',
                 reader_library='opencv', image_format='rgb', rgba_layout_color=0, test_mode=False, multilabel=False,
                 lazy_init=False, csv_path=None):
        annotation_path = annotation_path if annotation_path is not None else csv_path
        if annotation_path is None:
            raise ValueError('`annotation_path` must be specified.')
        super().__init__(transform, augment, input_dtype, reader_library, image_format, rgba_layout_color, test_mode)
        if num_classes is None and multilabel:
            raise ValueError('``num_classes`` must be specified when ``multilabel`` is `True`')
        self.data_folder, self.num_classes, self.multilabel, self.lazy_init = str(data_folder), num_classes, multilabel, lazy_init
        self.input_column, self.target_column, self.target_dtype = input_column, target_column, target_dtype
        self.df = _read_table(data_folder, annotation_path,
                              dtype={input_column: 'str', target_column: 'str' if multilabel else 'int'})
        if not lazy_init and not test_mode:
            self.df[target_column] = self.df[target_column].apply(self.process_function)

    def process_function(self, target):
        import re
        if self.multilabel:
            labels = list(map(int, re.findall(r'\d+', target)))
            if max(labels) >= self.num_classes:
                raise ValueError(f'Target column contains label: {max(labels)}, '
                                 f'it\'s more than num_classes = {self.num_classes}')
            multihot = np.zeros((self.num_classes,), dtype=bool)
            multihot[labels] = True
            return multihot
        if self.num_classes is not None and target >= self.num_classes:
            raise ValueError(f'Target column contains label: {target}, it\'s more than num_classes = {self.num_classes}')
        return target

    def get_raw(self, idx):
        record = self.df.iloc[idx]
        sample = {'image': _read_image(os.path.join(self.data_folder, record[self.input_column]), self.image_format,
                                       self.reader_library, self.rgba_layout_color), 'index': idx}
        if not self.test_mode:
            target = record[self.target_column]
            sample['target'] = self.process_function(target) if self.lazy_init else target
        return self._apply(self.augment, sample)

    def __getitem__(self, idx):
        sample = super().__getitem__(idx)
        if not self.test_mode:
            sample['target'] = torch.tensor(sample['target']).type(getattr(torch, self.target_dtype))
        return sample

    def __len__(self):
        return len(self.df)


@DATASETS.register_class
class ImageSegmentationDataset(ImageDataset):
    """csv / pkl of (image_path, mask_path) rows: torchok/data/datasets/segmentation/image_segmentation.py:14-125."""

    def __init__(self, data_folder, annotation_path, transform, augment=None, input_column='image_path',
                 input_dtype='float32', target_column='mask_path', target_dtype='int64', reader_library='opencv',
                 image_format='rgb', rgba_layout_color=0, test_mode=False):
        super().__init__(transform, augment, input_dtype, reader_library, image_format, rgba_layout_color, test_mode)
        self.data_folder, self.input_column, self.target_column = str(data_folder), input_column, target_column
        self.target_dtype = target_dtype
        self.df = _read_table(data_folder, annotation_path, dtype={input_column: 'str', target_column: 'str'})

    def get_raw(self, idx):
        record = self.df.iloc[idx]
        sample = {'image': _read_image(os.path.join(self.data_folder, record[self.input_column]), self.image_format,
                                       self.reader_library, self.rgba_layout_color), 'index': idx}
        if not self.test_mode:
            mask_path = os.path.join(self.data_folder, record[self.target_column])
            mask = cv2.imread(str(mask_path), 0)
            if mask is None:
                raise ValueError(f'{mask_path} was not read correctly!')
            sample['mask'] = mask
        return self._apply(self.augment, sample)

    def __getitem__(self, idx):
        sample = super().__getitem__(idx)
        if not self.test_mode:
            sample['target'] = torch.as_tensor(sample.pop('mask')).type(getattr(torch, self.target_dtype))
        return sample

    def __len__(self):
        return len(self.df)


@DATASETS.register_class
class SweetPepper(ImageSegmentationDataset):
    """torchok/data/datasets/examples/sweet_pepper.py:12-91: `<data_folder>/sweet_pepper/{train,valid}.csv`."""
    base_folder, train_csv, valid_csv = 'sweet_pepper', 'train.csv', 'valid.csv'

    def __init__(self, train, download, data_folder, transform, augment=None, input_dtype='float32',
                 target_dtype='int64', reader_library='opencv', image_format='rgb', rgba_layout_color=0,
                 test_mode=False):
        path = os.path.join(str(data_folder), self.base_folder)
        if not os.path.isdir(path):
            raise RuntimeError('Dataset not found or corrupted. You can use download=True to download it'
                               f' [torchok_b200: no network on this box; expected {path}]')
        super().__init__(path, self.train_csv if train else self.valid_csv, transform, augment, input_dtype=input_dtype,
                         target_column='mask', target_dtype=target_dtype, reader_library=reader_library,
                         image_format=image_format, rgba_layout_color=rgba_layout_color, test_mode=test_mode)


@DATASETS.register_class
class SOP(ImageDataset):
    """Stanford Online Products, torchok/data/datasets/examples/sop.py:15-136: `Ebay_{train,test}.txt` (space separated,
    columns class_id / path); targets are zero-based per split (train: class_id - 1, test: class_id - 11319)."""
    base_folder, train_txt, test_txt = 'Stanford_Online_Products', 'Ebay_train.txt', 'Ebay_test.txt'

    def __init__(self, train, download, data_folder, transform, augment=None, input_dtype='float32',
                 reader_library='opencv', image_format='rgb', rgba_layout_color=0, test_mode=False):
        super().__init__(transform, augment, input_dtype, reader_library, image_format, rgba_layout_color, test_mode)
        self.path, self.train = os.path.join(str(data_folder), self.base_folder), train
        if not os.path.isdir(self.path):
            raise RuntimeError('Dataset not found or corrupted. You can use download=True to download it'
                               f' [torchok_b200: no network on this box; expected {self.path}]')
        self.csv = _read_table(self.path, self.train_txt if train else self.test_txt, sep=' ')
        self.target_column, self.path_column = 'class_id', 'path'

    def get_raw(self, idx):
        record = self.csv.iloc[idx]
        sample = {'image': _read_image(os.path.join(self.path, record[self.path_column]), self.image_format,
                                       self.reader_library, self.rgba_layout_color), 'index': idx}
        if not self.test_mode:
            sample['target'] = record[self.target_column] - (1 if self.train else 11319)
        return self._apply(self.augment, sample)

    def __len__(self):
        return len(self.csv)


@TRANSFORMS.register_class
class FancyPCA(_Transform):
    """albumentations.FancyPCA (Krizhevsky et al. colour augmentation): add alpha_k * lambda_k * p_k summed over the
    principal components of the image's own RGB covariance, alpha_k ~ N(0, alpha)."""

    def __init__(self, alpha=0.1, always_apply=False, p=0.5):
        super().__init__(always_apply, p)
        self.alpha = alpha

    def params(self, sample):
        return {'a': np.random.normal(0.0, self.alpha, 3)}

    def image(self, img, a=None):
        if img.ndim != 3 or img.shape[2] != 3 or a is None:
            return img
        x = img.astype(np.float64).reshape(-1, 3) / 255.0
        x = x - x.mean(0)
        vals, vecs = np.linalg.eigh(np.cov(x, rowvar=False))
        order = vals.argsort()[::-1]
        vals, vecs = vals[order], vecs[:, order]
        shift = (vecs @ (a * vals).reshape(3, 1)).reshape(1, 1, 3) * 255.0
        return np.clip(img.astype(np.float64) + shift, 0, 255).astype(np.uint8)

    def mask(self, m, **params):
        return m
