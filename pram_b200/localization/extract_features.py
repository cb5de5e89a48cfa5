"""Offline local-feature export -- drop-in for the reference's ``localization/extract_features.py`` (``confs`` :26-86,
``ImageDataset`` :89-170, ``get_model`` :173-191, ``main`` :195-246; BASELINE.json config 1):

    python -m pram_b200.localization.extract_features --image_dir DIR --export_dir OUT --conf sfd2

For every image: SFD2 on the device through ``extract_sfd2_return`` (NMS radius 3, strict threshold, score order, top
``max_keypoints``), keypoints mapped back to the original resolution with the reference's ``(k + .5) * scale - .5`` rule,
and one group ``{keypoints [N,2] f64, scores [N] f64, descriptors [D,N] f64, image_size [2]}`` per image name in the
feature store (HDF5 through h5py when installed, else the ``.npz``-backed store of ``h5store.py`` with the same
interface).  The SuperPoint entries of the reference's table are outside the hot path and rejected.
"""
from __future__ import annotations

import argparse
import logging
import os
import os.path as osp
import pprint
from pathlib import Path
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

from ..nets.sfd2 import ResNet4x, extract_sfd2_return
from .h5store import open_store

_SFD2_MODEL = {
    'outdim': 128, 'name': 'resnet4x', 'use_stability': False, 'max_keypoints': 4096, 'conf_th': 0.005, 'multiscale': False,
    'scales': [1.0], 'model_fn': osp.join(os.getcwd(), 'weights/sfd2_20230511_210205_resnet4x.79.pth'),
}
# same keys and values as the reference table (extract_features.py:26-86); 'superpoint-n4096' is listed there too but
# SuperPoint is not part of the hot path
confs = {
    'resnet4x-20230511-210205-pho-0005': {
        'output': 'feats-resnet4x-20230511-210205-pho-0005', 'model': dict(_SFD2_MODEL),
        'preprocessing': {'grayscale': False, 'resize_max': False}, 'mask': False,
    },
    'sfd2': {
        'output': 'feats-sfd2', 'model': dict(_SFD2_MODEL),
        'preprocessing': {'grayscale': False, 'resize_max': False}, 'mask': False,
    },
}


class ImageDataset(torch.utils.data.Dataset):
    """Images under ``root`` (or listed in ``image_list``) as float32 CHW RGB in [0, 1] + name + original (w, h);
    reference extract_features.py:89-170."""
    default_conf = {'globs': ['*.jpg', '*.png', '*.jpeg', '*.JPG', '*.PNG'], 'grayscale': False, 'resize_max': None,
                    'resize_force': False}

    def __init__(self, root, conf, image_list=None, mask_root=None):
        self.conf = conf = SimpleNamespace(**{**self.default_conf, **conf})
        self.root = Path(root)
        self.mask_root = mask_root
        if image_list is None:
            paths = [p for g in conf.globs for p in self.root.glob('**/' + g)]
            if not paths:
                raise ValueError(f'Could not find any image in root: {root}.')
            self.paths = [p.relative_to(self.root) for p in paths]
        else:
            self.paths = [Path(l.strip()) for l in Path(image_list).read_text().splitlines() if l.strip()]
        logging.info(f'Found {len(self.paths)} images in root {root}.')

    def __len__(self):
        return len(self.paths)

    def __getitem__(self, idx):
        import cv2
        path = self.paths[idx]
        image = cv2.imread(str(self.root / path), cv2.IMREAD_GRAYSCALE if self.conf.grayscale else cv2.IMREAD_COLOR)
        if image is None:
            raise ValueError(f'Cannot read image {str(path)}.')
        if not self.conf.grayscale:
            image = image[:, :, ::-1]  # BGR -> RGB
        image = image.astype(np.float32)
        w, h = size = image.shape[:2][::-1]
        if self.conf.resize_max and (self.conf.resize_force or max(w, h) > self.conf.resize_max):
            scale = self.conf.resize_max / max(h, w)
            image = cv2.resize(image, (int(round(w * scale)), int(round(h * scale))), interpolation=cv2.INTER_CUBIC)
        image = image[None] if self.conf.grayscale else image.transpose((2, 0, 1))
        data = {'name': str(path), 'image': np.ascontiguousarray(image / 255.), 'original_size': np.array(size)}
        if self.mask_root is not None:
            mask_path = Path(str(path).replace('jpg', 'png'))
            if (Path(self.mask_root) / mask_path).exists():
                mask = cv2.imread(str(Path(self.mask_root) / mask_path))
                mask = cv2.resize(mask, dsize=(image.shape[2], image.shape[1]), interpolation=cv2.INTER_NEAREST)
            else:
                mask = np.zeros(shape=(image.shape[1], image.shape[2], 3), dtype=np.uint8)
            data['mask'] = mask
        return data


def get_model(model_name, weight_path, outdim=128, **kwargs):
    """-> (model, extractor); reference extract_features.py:173-191."""
    if model_name != 'resnet4x':
        raise ValueError(f'extractor {model_name!r} is not on the B200 hot path (only the SFD2 ResNet4x is)')
    model = ResNet4x(outdim=outdim).eval()
    model.load_state_dict(torch.load(weight_path, map_location='cpu', weights_only=False)['state_dict'], strict=True)
    return model, extract_sfd2_return


def export_one(pred: dict, image_shape, original_size) -> dict:
    """The reference's per-image post-processing (:226-236): descriptors to [D, N], keypoints to the original resolution."""
    pred = dict(pred)
    pred['descriptors'] = pred['descriptors'].transpose()
    pred['image_size'] = original_size = np.asarray(original_size)
    size = np.array(image_shape[-2:][::-1])
    scales = (original_size / size).astype(np.float32)
    pred['keypoints'] = (pred['keypoints'] + .5) * scales[None] - .5
    return pred


@torch.no_grad()
def main(conf, image_dir, export_dir, image_list: Optional[str] = None, device='cuda', overwrite: bool = False):
    logging.info('Extracting local features with configuration:' f'\n{pprint.pformat(conf)}')
    model, extractor = get_model(model_name=conf['model']['name'], weight_path=conf['model']['model_fn'],
                                 use_stability=conf['model']['use_stability'], outdim=conf['model']['outdim'])
    model = model.to(device)
    loader = torch.utils.data.DataLoader(ImageDataset(image_dir, conf['preprocessing'], image_list=image_list, mask_root=None),
                                         num_workers=0)
    os.makedirs(export_dir, exist_ok=True)
    feature_path = Path(export_dir, conf['output'] + '.h5')
    feature_file = open_store(feature_path, 'a')
    try:
        for data in loader:
            name = data['name'][0]
            if name in feature_file:
                if not overwrite:
                    continue
                del feature_file[name]
            pred = extractor(model, img=data['image'], topK=conf['model']['max_keypoints'], mask=None,
                             conf_th=conf['model']['conf_th'], scales=conf['model']['scales'])
            if not isinstance(pred, dict):  # no keypoint above the threshold
                continue
            pred = export_one(pred, tuple(data['image'].shape), data['original_size'][0].numpy())
            grp = feature_file.create_group(name)
            for k, v in pred.items():
                grp.create_dataset(k, data=v)
    finally:
        feature_file.close()
    logging.info('Finished exporting features.')
    return feature_path


if __name__ == '__main__':
    parser = argparse.ArgumentParser()
    parser.add_argument('--image_dir', type=Path, required=True)
    parser.add_argument('--image_list', type=str, default=None)
    parser.add_argument('--mask_dir', type=Path, default=None)
    parser.add_argument('--export_dir', type=Path, required=True)
    parser.add_argument('--conf', type=str, required=True, choices=list(confs.keys()))
    args = parser.parse_args()
    main(confs[args.conf], args.image_dir, args.export_dir, image_list=args.image_list)
