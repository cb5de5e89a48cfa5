"""Synthetic inputs, seeded random-init weights and checkpoint loading shared by bench.py, the tests and the oracle.

Nothing here computes any stage of the localization path: it only produces INPUTS (the seeded polygon frames of
SURVEY.md section 8d, state dicts with the reference's key schema, the reference's shipped checkpoints when they have
been staged) and the pose-error metric used to report how many frames localise.  It lives outside ``oracle/`` so
that bench.py's product arm does not touch the oracle at all.
"""
from __future__ import annotations

import math
import os
from pathlib import Path
from typing import Optional, Tuple

import numpy as np
import torch
from torch import Tensor

ROOT = Path(__file__).resolve().parent
REFERENCE_ROOT = Path(os.environ.get('PRAM_REFERENCE_ROOT', '/root/reference'))
STAGED = ROOT / 'oracle' / '_ref' / 'weights'   # reference artefacts staged for the GPU box (git-ignored)

SFD2_WEIGHT = 'sfd2_20230511_210205_resnet4x.79.pth'
GML_WEIGHT = 'imp_gml.920.pth'

RGB_MEAN = (0.485, 0.456, 0.406)  # reference nets/sfd2.py:14
RGB_STD = (0.229, 0.224, 0.225)  # reference nets/sfd2.py:15


def weight_path(name: str) -> Optional[Path]:
    for root in (REFERENCE_ROOT / 'weights', STAGED):
        p = root / name
        if p.exists():
            return p
    return None


def _torch_load(path: Path):
    """torch>=2.6 defaults to weights_only=True and the GML checkpoint holds a numpy scalar
    (SURVEY.md section 5); these files are the reference's own artefacts, so load them fully."""
    return torch.load(str(path), map_location='cpu', weights_only=False)


def load_sfd2_state() -> Optional[dict]:
    p = weight_path(SFD2_WEIGHT)
    return None if p is None else _torch_load(p)['state_dict']


def load_gml_state() -> Optional[dict]:
    p = weight_path(GML_WEIGHT)
    return None if p is None else _torch_load(p)['model']


# ---- seeded random state dicts (no shipped weights exist for SegNetViT / AdaGML) --------------

def _lin(g, out_f, in_f, scale=1.0):
    bound = 1.0 / np.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound * scale
    b = (torch.rand(out_f, generator=g) * 2 - 1) * bound
    return w, b


def _mlp_state(sd, g, pre, d_in, d_hid, d_out):
    sd[pre + '.0.weight'], sd[pre + '.0.bias'] = _lin(g, d_hid, d_in)
    sd[pre + '.1.weight'] = 1 + 0.1 * torch.randn(d_hid, generator=g)
    sd[pre + '.1.bias'] = 0.1 * torch.randn(d_hid, generator=g)
    sd[pre + '.3.weight'], sd[pre + '.3.bias'] = _lin(g, d_out, d_hid)


def random_segnetvit_state(n_class=113, n_layers=15, output_dim=1024, desc_dim=256, seed=0) -> dict:
    """Seeded state dict with the reference's SegNetViT key schema."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for i in range(n_layers):
        p = f'gnn.layers.{i}'
        sd[p + '.qkv.weight'], sd[p + '.qkv.bias'] = _lin(g, 768, 256)
        sd[p + '.proj.weight'], sd[p + '.proj.bias'] = _lin(g, 256, 256)
        _mlp_state(sd, g, p + '.mlp', 512, 512, 256)
    sd['kenc.Wr.weight'] = torch.randn(32, 2, generator=g)
    sd['input_proj.weight'], sd['input_proj.bias'] = _lin(g, 256, desc_dim)
    _mlp_state(sd, g, 'seg', 256, output_dim, n_class)
    return sd


def random_gml_state(n_layers=9, seed=0, ada=False) -> dict:
    """Seeded state dict with the reference's GML / AdaGML key schema."""
    g = torch.Generator().manual_seed(seed)
    sd = {'bin_score': torch.tensor(1.0)}
    sd['input_proj.weight'], sd['input_proj.bias'] = _lin(g, 256, 128)
    sd['poseenc.Wr.weight'] = torch.randn(32, 2, generator=g)
    for i in range(n_layers):
        p = f'self_attn.{i}'
        sd[p + '.qkv.weight'], sd[p + '.qkv.bias'] = _lin(g, 768, 256)
        sd[p + '.proj.weight'], sd[p + '.proj.bias'] = _lin(g, 256, 256)
        _mlp_state(sd, g, p + '.mlp', 512, 512, 256)
        p = f'cross_attn.{i}'
        sd[p + '.to_qk.weight'], sd[p + '.to_qk.bias'] = _lin(g, 256, 256)
        sd[p + '.to_v.weight'], sd[p + '.to_v.bias'] = _lin(g, 256, 256)
        sd[p + '.proj.weight'], sd[p + '.proj.bias'] = _lin(g, 256, 256)
        _mlp_state(sd, g, p + '.mlp', 512, 512, 256)
        sd[f'out_proj.{i}.weight'], sd[f'out_proj.{i}.bias'] = _lin(g, 256, 256)
        if ada:
            p = f'pooling.{i}'
            _mlp_state(sd, g, p + '.score_enc', 2, 256, 256)
            sd[p + '.proj.weight'], sd[p + '.proj.bias'] = _lin(g, 256, 256)
            _mlp_state(sd, g, p + '.predict', 512, 256, 1)
    return sd


def random_sfd2_state(seed=0) -> dict:
    """Seeded SFD2 state dict (reference key schema) with non-trivial BN statistics, for
    conv-stack parity when the shipped checkpoint is not available."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, co, ci, k, bias=True, groups=1):
        fan = ci // groups * k * k
        sd[name + '.weight'] = torch.randn(co, ci // groups, k, k, generator=g) * np.sqrt(2.0 / fan)
        if bias:
            sd[name + '.bias'] = 0.1 * torch.randn(co, generator=g)

    def bn(name, c):
        sd[name + '.weight'] = 1 + 0.1 * torch.randn(c, generator=g)
        sd[name + '.bias'] = 0.1 * torch.randn(c, generator=g)
        sd[name + '.running_mean'] = 0.1 * torch.randn(c, generator=g)
        sd[name + '.running_var'] = 0.5 + torch.rand(c, generator=g)
        sd[name + '.num_batches_tracked'] = torch.tensor(0)

    for name, ci, co in (('conv1a', 3, 64), ('conv1b', 64, 64), ('conv2a', 64, 128),
                         ('conv2b', 128, 128), ('conv3a', 128, 256), ('conv3b', 256, 256)):
        conv(name + '.0', co, ci, 3)
        bn(name + '.1', co)
    for i in range(3):
        p = f'conv4.{i}'
        conv(p + '.conv1', 256, 256, 1, bias=False)
        bn(p + '.bn1', 256)
        conv(p + '.conv2', 256, 256, 3, bias=False, groups=32)
        bn(p + '.bn2', 256)
        conv(p + '.conv3', 256, 256, 1, bias=False)
        bn(p + '.bn3', 256)
    for head in ('convPa', 'convDa'):
        conv(head + '.0', 256, 256, 3)
        bn(head + '.1', 256)
        conv(head + '.3', 256, 256, 3)
    conv('convPb', 65, 256, 1)
    conv('convDb', 128, 256, 1)
    return sd


def calibrated_adagml_state(seed: int = 7, gain: float = 10.0, bias: float = 0.8) -> dict:
    """Seeded AdaGML state whose pooling confidences straddle the pruning thresholds (default init sits at
    ~0.5 < 0.56 and collapses every token set): the last pooling layer is scaled / biased so
    that tokens are pruned over several layers before the early exit fires."""
    sd = random_gml_state(seed=seed, ada=True)
    for i in range(9):
        sd[f'pooling.{i}.predict.3.weight'] = sd[f'pooling.{i}.predict.3.weight'] * gain
        sd[f'pooling.{i}.predict.3.bias'] = torch.full((1,), bias)
    return sd


# ---- synthetic frames -- seeded, dataset-free -------------------------------------------

def polys_frame(h: int = 480, w: int = 640, seed: int = 0, n_poly: int = 600) -> np.ndarray:
    """Grey canvas + random filled polygons + sigma=0.8 blur -> float32 RGB [h,w,3] in [0,1]."""
    import cv2
    rs = np.random.RandomState(seed)
    img = np.full((h, w, 3), 127, np.uint8)
    for _ in range(n_poly):
        k = rs.randint(3, 7)
        off = np.array([rs.randint(0, w), rs.randint(0, h)])
        pts = (rs.randint(0, 60, size=(k, 2)) + off - 30).astype(np.int32)
        col = tuple(int(c) for c in rs.randint(0, 256, size=3))
        cv2.fillPoly(img, [pts], col)
    img = cv2.GaussianBlur(img, (0, 0), 0.8)
    return img.astype(np.float32) / 255.0


def frame_tensor(h: int = 480, w: int = 640, seed: int = 0) -> Tensor:
    """ImageNet-normalised [1,3,h,w] tensor of ``polys_frame`` (the online loop's preprocessing,
    reference localization/loc_by_rec_online.py:86-106)."""
    img = torch.from_numpy(polys_frame(h, w, seed)).permute(2, 0, 1)[None]
    mean = torch.tensor(RGB_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(RGB_STD).view(1, 3, 1, 1)
    return ((img - mean) / std).contiguous()


# ---- pose metric ------------------------------------------------------------------------------------------------

def quat_to_rotmat(q: np.ndarray) -> np.ndarray:
    """wxyz quaternion -> R; same convention as reference colmap_utils/read_write_model.py:556."""
    w, x, y, z = q
    return np.array([
        [1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * w * z, 2 * z * x + 2 * w * y],
        [2 * x * y + 2 * w * z, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * w * x],
        [2 * z * x - 2 * w * y, 2 * y * z + 2 * w * x, 1 - 2 * x * x - 2 * y * y]])


def rotmat_to_quat(R: np.ndarray) -> np.ndarray:
    """R -> wxyz quaternion with w >= 0."""
    K = np.array([
        [R[0, 0] - R[1, 1] - R[2, 2], 0, 0, 0],
        [R[0, 1] + R[1, 0], R[1, 1] - R[0, 0] - R[2, 2], 0, 0],
        [R[0, 2] + R[2, 0], R[1, 2] + R[2, 1], R[2, 2] - R[0, 0] - R[1, 1], 0],
        [R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1], R[0, 0] + R[1, 1] + R[2, 2]]]) / 3.0
    vals, vecs = np.linalg.eigh(K)
    q = vecs[[3, 0, 1, 2], np.argmax(vals)]
    return -q if q[0] < 0 else q


def pose_error(q_pred, t_pred, q_gt, t_gt) -> Tuple[float, float]:
    """(rotation error in degrees, camera-centre error); reference localization/utils.py:30-53."""
    Rp, Rg = quat_to_rotmat(np.asarray(q_pred, float)), quat_to_rotmat(np.asarray(q_gt, float))
    cp = -Rp.T @ np.asarray(t_pred, float).reshape(3)
    cg = -Rg.T @ np.asarray(t_gt, float).reshape(3)
    d = min(1.0, max(-1.0, abs(float(np.dot(q_pred, q_gt)))))
    return 2 * math.acos(d) * 180 / math.pi, float(np.linalg.norm(cp - cg))
