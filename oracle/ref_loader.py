"""Locate the reference's *artefacts* (checkpoints) and, in the build container only, import the
reference's own Python modules to pin the oracle.  TEST INFRASTRUCTURE ONLY (see pram_oracle.py).

* ``/root/reference`` exists only in the build container.  ``stage_weights()`` (called from
  ``__graft_entry__.build()``) copies the two shipped checkpoints into ``oracle/_ref/weights/``,
  which is git-ignored (never enters history; licence CC BY-NC 4.0, reference LICENSE:1) but not
  gpurun-ignored, so the GPU box can run weight-dependent parity tests.  No reference *source* is
  ever copied.
* ``import_reference()`` puts ``/root/reference`` on ``sys.path`` and applies the one CPU shim the
  survey lists (nets.adagml.sink_algorithm hard-codes 'cuda', reference nets/adagml.py:45-48).
"""
from __future__ import annotations

import os
import shutil
import sys
from pathlib import Path
from typing import Optional

import numpy as np
import torch

REFERENCE_ROOT = Path(os.environ.get('PRAM_REFERENCE_ROOT', '/root/reference'))
HERE = Path(__file__).resolve().parent
STAGED = HERE / '_ref' / 'weights'

SFD2_WEIGHT = 'sfd2_20230511_210205_resnet4x.79.pth'
GML_WEIGHT = 'imp_gml.920.pth'


def reference_available() -> bool:
    return (REFERENCE_ROOT / 'nets' / 'sfd2.py').exists()


def stage_weights() -> None:
    if not reference_available():
        return
    STAGED.mkdir(parents=True, exist_ok=True)
    for name in (SFD2_WEIGHT, GML_WEIGHT):
        src, dst = REFERENCE_ROOT / 'weights' / name, STAGED / name
        if src.exists() and (not dst.exists() or dst.stat().st_size != src.stat().st_size):
            shutil.copyfile(src, dst)


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


def import_reference():
    """Returns a namespace of the reference's own modules (build container only)."""
    if not reference_available():
        raise RuntimeError('reference tree not mounted')
    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))
    import nets.sfd2 as ref_sfd2
    import nets.segnetvit as ref_segnetvit
    import nets.gml as ref_gml
    import nets.adagml as ref_adagml
    import nets.utils as ref_utils
    ref_adagml.sink_algorithm = ref_gml.sink_algorithm  # device-agnostic, identical maths
    from types import SimpleNamespace
    return SimpleNamespace(sfd2=ref_sfd2, segnetvit=ref_segnetvit, gml=ref_gml, adagml=ref_adagml,
                           utils=ref_utils)


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
