"""Batched feature matching -- the device side of reference ``localization/match_features_batch.py``.

``confs`` keeps the reference's names and model configs for the matchers that are on the hot path (``gml``,
``adagml``, ``NNM``; reference match_features_batch.py:17-61).  The reference streams pairs one at a time from an
h5 file through a DataLoader (batch_size 1) and writes ``matches0`` (int16) / ``matching_scores0`` (fp16) per pair
(``writer_fn``, :119-129).  Here pairs with equal keypoint counts are stacked and matched in ONE call (the matcher
kernels are batched over pairs -- this is the "multi-landmark match_features_batch" of BASELINE.json config 5), and
the records are returned / written in the reference's dtypes.  Feature stores are h5-like mappings
(``store[name]['keypoints' | 'descriptors' ([D, N]) | 'scores' | 'image_size'][()]``): h5py files when h5py is
installed, plain dicts of numpy arrays otherwise.
"""
from __future__ import annotations

from collections import defaultdict
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Tuple, Union

import numpy as np
import torch

from . import matchers
from .base_model import dynamic_load

confs = {
    'gml': {
        'output': 'gml',
        'model': {'name': 'gml', 'weight_path': 'weights/imp_gml.920.pth', 'sinkhorn_iterations': 20},
    },
    'adagml': {
        'output': 'adagml',
        'model': {'name': 'adagml', 'weight_path': 'weights/imp_adagml.80.pth', 'sinkhorn_iterations': 20},
    },
    'NNM': {
        'output': 'NNM',
        'model': {'name': 'nearest_neighbor', 'do_mutual_check': True, 'distance_threshold': None},
    },
}


def names_to_pair(name0: str, name1: str, separator: str = '/') -> str:
    """reference colmap_utils/parsers.py: names_to_pair."""
    return separator.join((name0.replace('/', '-'), name1.replace('/', '-')))


def load_matcher(conf: Dict, device='cuda'):
    Model = dynamic_load(matchers, conf['model']['name'])
    return Model(conf['model']).eval().to(device)


def _read(store, name: str) -> Dict[str, np.ndarray]:
    grp = store[name]
    return {k: np.asarray(grp[k][()]) for k in ('keypoints', 'descriptors', 'scores', 'image_size')}


def _pack(feats: List[Dict[str, np.ndarray]], suffix: str, device, nn_layout: bool) -> Dict[str, torch.Tensor]:
    """Stack equally sized feature sets; descriptors are stored [D, N] (reference extract_features.py:223-238) and
    the attentional matchers take [B, N, D] (FeaturePairsDataset transposes, :98-99), the NN matcher [B, D, N]."""
    t = lambda key: torch.from_numpy(np.stack([np.ascontiguousarray(f[key]) for f in feats])).float().to(device)
    desc = t('descriptors')
    w, h = (int(v) for v in feats[0]['image_size'][:2])
    return {'keypoints' + suffix: t('keypoints'), 'scores' + suffix: t('scores'),
            'descriptors' + suffix: desc if nn_layout else desc.transpose(1, 2).contiguous(),
            'image' + suffix: torch.empty((1, 1, h, w), device='meta')}


@torch.no_grad()
def match_pairs(conf: Dict, pairs: Iterable[Tuple[str, str]], features_q, features_ref=None, model=None, device='cuda',
                max_batch: int = 16) -> Dict[str, Dict[str, np.ndarray]]:
    """-> {names_to_pair(name0, name1): {'matches0': int16 [N0], 'matching_scores0': float16 [N0]}}."""
    features_ref = features_q if features_ref is None else features_ref
    model = model if model is not None else load_matcher(conf, device)
    nn_layout = conf['model']['name'] == 'nearest_neighbor'
    batchable = conf['model']['name'] != 'adagml'  # AdaGML prunes per pair (reference adagml.py:358: B = 1 only)
    groups = defaultdict(list)
    cache: Dict[Tuple[int, str], Dict[str, np.ndarray]] = {}

    def get(store, which, name):
        key = (which, name)
        if key not in cache:
            cache[key] = _read(store, name)
        return cache[key]
    for name0, name1 in pairs:
        f0, f1 = get(features_q, 0, name0), get(features_ref, 1, name1)
        shape_key = (f0['keypoints'].shape[0], f1['keypoints'].shape[0], tuple(f0['image_size']), tuple(f1['image_size']))
        groups[shape_key if batchable else (name0, name1)].append((name0, name1, f0, f1))
    out: Dict[str, Dict[str, np.ndarray]] = {}
    for items in groups.values():
        for i in range(0, len(items), max_batch):
            chunk = items[i:i + max_batch]
            data = {**_pack([c[2] for c in chunk], '0', device, nn_layout), **_pack([c[3] for c in chunk], '1', device, nn_layout)}
            pred = model(data)
            m = pred['matches0'].cpu().short().numpy()
            s = pred['matching_scores0'].cpu().half().numpy() if 'matching_scores0' in pred else None
            for j, (name0, name1, _, _) in enumerate(chunk):
                rec = {'matches0': m[j]}
                if s is not None:
                    rec['matching_scores0'] = s[j]
                out[names_to_pair(name0, name1)] = rec
    return out


def main(conf: Dict, pairs: Union[Path, List[Tuple[str, str]]], features, export_dir: Optional[Path] = None,
         matches: Optional[Path] = None, features_ref=None, overwrite: bool = False):
    """Reference entry point (match_features_batch.py:132-178).  ``pairs`` is a pairs file ("name0 name1" per line) or
    a list; ``features`` / ``features_ref`` are store paths (``h5store.open_store``) or in-memory mappings.  Returns the matches path
    when an h5 file is written, else the dict of records."""
    if isinstance(pairs, (str, Path)):
        pairs = [tuple(l.split()[:2]) for l in Path(pairs).read_text().splitlines() if l.strip()]
    opened = []

    def open_store(f):
        if isinstance(f, (str, Path)):
            from .h5store import open_store as _open
            fd = _open(f, 'r')  # HDF5 through h5py when installed, else the .npz-backed store (same interface)
            opened.append(fd)
            return fd
        return f
    fq = open_store(features)
    fr = open_store(features_ref) if features_ref is not None else None
    try:
        recs = match_pairs(conf, pairs, fq, fr)
    finally:
        for fd in opened:
            fd.close()
    if matches is None and export_dir is not None and isinstance(features, (str, Path)):
        matches = Path(export_dir, f'{Path(features).stem}-{conf["output"]}.h5')
    if matches is None:
        return recs
    from .h5store import open_store as _open
    Path(matches).parent.mkdir(parents=True, exist_ok=True)
    fd = _open(matches, 'a', libver='latest')
    try:
        for pair, rec in recs.items():
            if pair in fd:
                if not overwrite:
                    continue
                del fd[pair]
            grp = fd.create_group(pair)
            for k, v in rec.items():
                grp.create_dataset(k, data=v)
    finally:
        fd.close()
    return matches
