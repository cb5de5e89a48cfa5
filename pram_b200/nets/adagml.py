"""B200-native AdaGML matcher -- drop-in for reference ``nets/adagml.py``.

GML plus a per-layer ``PoolingLayer`` confidence used for data-dependent token pruning (layers >= 1,
sets with >= n_min_tokens tokens) and early exit (> 95 % confident), reference nets/adagml.py:307-404.
The control flow is data dependent, so -- exactly like the reference -- one small device->host read per
layer decides pruning / stopping and the batch size must be 1.  All tensor arithmetic (attention with
the per-token mean attention, pooling MLPs, distance, Sinkhorn, matches) runs in libpram_b200 kernels;
torch is used for the boolean-mask gathers of the surviving tokens and the final scatter.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from .. import _lib, ops
from . import _blocks as B
from .gml import GML


class PoolingLayer(nn.Module):
    """Parameter holder (reference nets/adagml.py:114-130)."""

    def __init__(self, hidden_dim: int = 256, score_dim: int = 2):
        super().__init__()
        self.score_enc = B.mlp_holder(score_dim, hidden_dim, hidden_dim)
        self.proj = nn.Linear(hidden_dim, hidden_dim)
        self.predict = B.mlp_holder(hidden_dim * 2, hidden_dim, 1)


class AdaGML(GML):
    default_config = {**GML.default_config, 'with_pose': True, 'min_confidence': 0.9,
                      'classification_background_weight': 0.05, 'pretrained': True}

    def __init__(self, config: Optional[dict] = None):
        super().__init__(config)
        self.n_min_tokens = self.config['n_min_tokens']
        self.pooling = nn.ModuleList([PoolingLayer(256, 2) for _ in range(self.n_layers)])

    def prepare(self):
        if self._packed is None:
            pk = super().prepare()
            pk['pool'] = [{**B.pack_mlp(pl.score_enc, 'se'), 'proj.w': B._c(pl.proj.weight),
                           'proj.b': B._c(pl.proj.bias), **B.pack_mlp(pl.predict, 'pr')} for pl in self.pooling]
        return self._packed

    def confidence_threshold(self, layer_index: int) -> float:
        """0.5 + 0.1*exp(-4 l / L), reference nets/adagml.py:516-520."""
        return float(np.clip(0.5 + 0.1 * np.exp(-4.0 * layer_index / self.n_layers), 0, 1))

    @staticmethod
    def _confidence(pp, x: torch.Tensor, ldx: int, att: torch.Tensor) -> torch.Tensor:
        """PoolingLayer.forward (reference nets/adagml.py:132-138) -> pre-sigmoid logits [T]."""
        T = att.shape[0]
        dev = att.device
        cat = torch.empty((T, 512), device=dev, dtype=torch.float32)
        hid = torch.empty((T, 256), device=dev, dtype=torch.float32)
        ops.linear_f32(x, ldx, pp['proj.w'], pp['proj.b'], cat, 512, T, 256, 256)
        B.run_mlp(pp, 'se', att, 2, T, 2, 256, 256, hid, cat[:, 256:], 512)
        z = torch.empty((T, 1), device=dev, dtype=torch.float32)
        B.run_mlp(pp, 'pr', cat, 512, T, 512, 256, 1, hid, z, 1)
        return z[:, 0]

    def forward(self, data, mode=0):
        if self.training:
            raise _lib.PramError('training is out of scope; call .eval()')
        if mode != 0:
            raise _lib.PramError('AdaGML.run (mode=1, training-time evaluation helper) is out of scope')
        return self.produce_matches(data)

    @torch.no_grad()
    def produce_matches(self, data: Dict[str, torch.Tensor], p: float = 0.2, **kwargs):
        d0, d1 = data['descriptors0'], data['descriptors1']
        _lib.require_cuda(d0, 'descriptors0')
        if d0.shape[0] != 1:
            raise ValueError('AdaGML prunes tokens with boolean masks and requires batch size 1 '
                             '(reference nets/adagml.py:358)')
        pk = self.prepare()
        dev = d0.device
        m_full, n_full = d0.shape[1], d1.shape[1]
        cos, sin = self._encode(pk, data)
        ws = B.Workspace(m_full + n_full, dev, self._split)
        ws.keep_f32 = True  # the pooling MLPs and the token compaction read the fp32 activation rows
        self._input_tokens(pk, ws, d0, d1)
        ind0 = torch.arange(m_full, device=dev)
        ind1 = torch.arange(n_full, device=dev)
        m, n = m_full, n_full
        ni = 0
        self.last_trace = []  # (layer, tokens0, tokens1) after pruning, for tests / diagnostics
        for ni in range(self.n_layers):
            seg0, seg1 = (0, 1, m), (m, 1, n)
            att = torch.empty((m + n, 2), device=dev, dtype=torch.float32)
            a_self = torch.empty((m + n,), device=dev, dtype=torch.float32)
            a_cross = torch.empty((m + n,), device=dev, dtype=torch.float32)
            B.self_block(ws, pk['self'][ni], (seg0, seg1), cos, sin, colmeans=[a_self[:m], a_self[m:]])
            B.cross_block(ws, pk['cross'][ni], seg0, seg1, colmeans=[a_cross[:m], a_cross[m:]])
            att[:, 0], att[:, 1] = a_self, a_cross
            conf = torch.sigmoid(self._confidence(pk['pool'][ni], ws.x, 512, att))
            if ni >= 1:
                th = self.confidence_threshold(ni)
                keep = conf > th
                k0, k1 = keep[:m], keep[m:]
                if m < self.n_min_tokens:
                    k0 = torch.ones_like(k0)
                if n < self.n_min_tokens:
                    k1 = torch.ones_like(k1)
                stop = bool((1.0 - (conf < th).float().sum() / (m_full + n_full)) > 0.95)  # host read
                sel = torch.cat([k0, k1])
                new_m, new_n = int(k0.sum()), int(k1.sum())
                if new_m != m or new_n != n:
                    ind0, ind1 = ind0[k0], ind1[k1]
                    x = ws.x[:, :256][sel].contiguous()
                    cos, sin = cos[sel].contiguous(), sin[sel].contiguous()
                    m, n = new_m, new_n
                    if m == 0 or n == 0:
                        raise ValueError('AdaGML pruned a keypoint set to zero tokens (the reference raises at '
                                         'nets/adagml.py:500 in the same situation)')
                    ws = B.Workspace(m + n, dev, self._split)
                    ws.keep_f32 = True
                    ws.x[:, :256] = x
                    if ws.split:
                        xs = ops.split_bf16(x, ws.split == 3)
                        ws.x_bf.hi[:, :256] = xs.hi
                        if xs.lo is not None:
                            ws.x_bf.lo[:, :256] = xs.lo
                self.last_trace.append((ni, m, n))
                if stop:
                    break
        dist = self._distance(pk, ni, ws, 1, m, n)
        i0, _, s0, _ = ops.sinkhorn_match(dist, pk['bin'], self.sinkhorn_iterations, p)
        valid = i0[0] > -1
        full_i = torch.full((1, m_full), -1, device=dev, dtype=torch.int64)
        full_i[0, ind0[valid]] = ind1[i0[0][valid]]
        full_s = torch.zeros((1, m_full), device=dev, dtype=torch.float32)
        full_s[0, ind0] = s0[0]
        return {'matches0': full_i, 'matching_scores0': full_s}
