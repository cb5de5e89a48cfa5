"""B200-native AdaGML matcher -- drop-in for reference ``nets/adagml.py``.

GML plus a per-layer ``PoolingLayer`` confidence used for data-dependent token pruning (layers >= 1,
sets with >= n_min_tokens tokens) and early exit (> 95 % confident), reference nets/adagml.py:307-404.

Tensor-core precisions (``bf16x3`` / ``bf16``): the whole control flow runs ON THE DEVICE (csrc/adagml_ops.cu) -- per-pair
token counts, a stable compaction kernel instead of boolean-mask indexing, the stop test as a per-pair flag whose state
is latched at the layer a pair exits, mean attention per token from the tcgen05 attention kernel's row statistics -- so
there is no host read inside the layer loop (the call can be captured into a CUDA graph) and, as an extension of the
reference, a BATCH of pairs can run together (``produce_matches_batched``; ``produce_matches`` keeps the reference's
batch-1 contract).  ``fp32`` precision (the exact-arithmetic CUDA-core mode) keeps the reference's structure: one small
device->host read per layer decides pruning / stopping and the surviving tokens are gathered with boolean masks.
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
        self._host_trace, self._device_state = None, None

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
    def produce_matches_batched(self, data: Dict[str, torch.Tensor], p: float = 0.2, check: bool = False):
        """Device-side AdaGML over a batch of pairs (tensor-core precisions).  Same dict as ``produce_matches`` with
        [B, M | N, .] tensors (+ optional ``num_keypoints0/1`` [B] for padded sets); no host synchronisation unless
        ``check`` (then a set pruned to zero tokens raises like the reference).  Returns ``matches0``,
        ``matching_scores0`` [B, M] plus the exit state (``stop_layer`` [B], ``num_tokens0/1`` [B], ``token_trace``
        [layers, 2, B] = tokens per set after each layer's pruning)."""
        d0, d1 = data['descriptors0'], data['descriptors1']
        _lib.require_cuda(d0, 'descriptors0')
        if not self._split:
            raise _lib.PramError('the device-side AdaGML path needs a tensor-core precision (bf16x3 / bf16)')
        pk = self.prepare()
        dev = d0.device
        b, m, _ = d0.shape
        n = d1.shape[1]
        if m == 0 or n == 0:
            raise ValueError('AdaGML needs at least one keypoint per set')
        L, D = self.n_layers, B.D
        T = b * (m + n)
        cos, sin = self._encode(pk, data)
        ws = self._workspace(T, dev)
        ws.keep_f32 = True  # the pooling MLPs read the fp32 activation rows
        self._input_tokens(pk, ws, d0, d1)
        seg0, seg1 = (0, b, m), (b * m, b, n)
        full = lambda v: torch.full((b,), v, device=dev, dtype=torch.int32)
        c0, c1 = self._counts(data.get('num_keypoints0'), b, dev), self._counts(data.get('num_keypoints1'), b, dev)
        cnt0 = c0.clamp(max=m).clone() if c0 is not None else full(m)   # updated in place by the prune kernel
        cnt1 = c1.clamp(max=n).clone() if c1 is not None else full(n)
        full0, full1 = cnt0.clone(), cnt1.clone()   # the pairs' original sizes: denominator of the stop test
        ind = torch.cat([torch.arange(m, device=dev, dtype=torch.int32).repeat(b),
                         torch.arange(n, device=dev, dtype=torch.int32).repeat(b)])
        ind2, cos2, sin2 = torch.empty_like(ind), torch.empty_like(cos), torch.empty_like(sin)
        # active[l] = pairs still running at the START of layer l (the launch predicate of that layer's kernels)
        stop_layer, active = full(-1), torch.full((L,), b, device=dev, dtype=torch.int32)
        err = torch.zeros((1,), device=dev, dtype=torch.int32)
        trace = torch.zeros((L, 2, b), device=dev, dtype=torch.int32)
        fcnt0, fcnt1 = torch.empty_like(cnt0), torch.empty_like(cnt1)
        dest = torch.empty((T,), device=dev, dtype=torch.int32)
        ind_fin = torch.empty_like(ind)
        lo = ws.split == 3
        md_tmp, md_fin = ops.empty_split((T, D), dev, lo), ops.empty_split((T, D), dev, lo)
        att = torch.empty((T, 2), device=dev, dtype=torch.float32)
        early_exit = self.config.get('device_early_exit', True)
        try:
            for ni in range(L):
                if ni >= 2 and early_exit:
                    # from here on every pair may have exited: the kernels of layer ni are launched with active[ni] as their
                    # predicate and return at once when it is 0 -- the reference's `break` (nets/adagml.py:370-372)
                    _lib.call('pram_set_launch_predicate', _lib.ptr(active[ni:]))
                counts = [cnt0, cnt1]
                B.self_block(ws, pk['self'][ni], (seg0, seg1), cos, sin, colmeans=[att[:b * m, 0], att[b * m:, 0]],
                             counts=counts)
                B.cross_block(ws, pk['cross'][ni], seg0, seg1, colmeans=[att[:b * m, 1], att[b * m:, 1]], counts=counts)
                z = self._confidence(pk['pool'][ni], ws.x, 2 * D, att)
                do_prune, last = ni >= 1, ni == L - 1
                if not (do_prune or last):
                    continue
                th = float(np.float32(self.confidence_threshold(ni)))
                _lib.call('pram_adagml_prune', _lib.ptr(z), b, m, n, _lib.ptr(cnt0), _lib.ptr(cnt1), th, int(self.n_min_tokens),
                          int(do_prune), ni, L, _lib.ptr(dest), _lib.ptr(stop_layer), _lib.ptr(active), _lib.ptr(trace),
                          _lib.ptr(err), _lib.ptr(full0), _lib.ptr(full1), _lib.stream_ptr())
                if do_prune:
                    src, dst = ws.cat_bf[ws.cur], ws.cat_bf[ws.cur ^ 1]
                    _lib.call('pram_adagml_move', _lib.ptr(dest), b, m, n, _lib.ptr(src.hi), _lib.ptr(src.lo), _lib.ptr(dst.hi),
                              _lib.ptr(dst.lo), 2 * D, _lib.ptr(ws.cat[ws.cur]), _lib.ptr(ws.cat[ws.cur ^ 1]), 2 * D,
                              _lib.ptr(cos), _lib.ptr(sin), _lib.ptr(cos2), _lib.ptr(sin2), _lib.ptr(ind), _lib.ptr(ind2),
                              _lib.stream_ptr())
                    ws.cur ^= 1
                    cos, cos2, sin, sin2, ind, ind2 = cos2, cos, sin2, sin, ind2, ind
                # the state a pair exits with: out_proj[ni] of its (pruned) tokens, index maps, counts
                ops.linear_tc(ws.x_bf, 2 * D, T, D, pk['out.tc'][ni], D, pk['out'][ni][1], out_bf=md_tmp, ld_bf=D,
                              split=ws.split)
                _lib.call('pram_adagml_latch', _lib.ptr(stop_layer), ni, b, m, n, _lib.ptr(md_tmp.hi), _lib.ptr(md_tmp.lo),
                          _lib.ptr(md_fin.hi), _lib.ptr(md_fin.lo), _lib.ptr(ind), _lib.ptr(ind_fin), _lib.ptr(cnt0),
                          _lib.ptr(cnt1), _lib.ptr(fcnt0), _lib.ptr(fcnt1), _lib.stream_ptr())
        finally:
            _lib.call('pram_set_launch_predicate', None)
        dist = torch.empty((b, m, n), device=dev, dtype=torch.float32)
        ops.linear_tc(md_fin, D, m, D, ops.split_rows(md_fin, b * m), n, out_f32=dist, ld_f32=n, split=ws.split, batch=b,
                      w_batched=True)
        i0, _, s0, _ = ops.sinkhorn_match(dist, pk['bin'], self.sinkhorn_iterations, p, m_counts=fcnt0, n_counts=fcnt1)
        full_i = torch.empty((b, m), device=dev, dtype=torch.int64)
        full_s = torch.empty((b, m), device=dev, dtype=torch.float32)
        _lib.call('pram_adagml_scatter', _lib.ptr(i0), _lib.ptr(s0), _lib.ptr(ind_fin), _lib.ptr(fcnt0), _lib.ptr(fcnt1), b, m, n,
                  _lib.ptr(full_i), _lib.ptr(full_s), _lib.stream_ptr())
        self._device_state = (stop_layer, trace)
        self._host_trace = None
        if check and int(err.item()):
            raise ValueError('AdaGML pruned a keypoint set to zero tokens (the reference raises at nets/adagml.py:500 in the '
                             'same situation)')
        return {'matches0': full_i, 'matching_scores0': full_s, 'stop_layer': stop_layer, 'num_tokens0': fcnt0,
                'num_tokens1': fcnt1, 'token_trace': trace}

    @property
    def last_trace(self):
        """[(layer, tokens0, tokens1)] after each pruning step of pair 0 of the last call (tests / diagnostics; on the
        device path this is where the only device->host read happens)."""
        if self._host_trace is None and self._device_state is not None:
            stop_layer, trace = self._device_state
            sl, tr = int(stop_layer[0].item()), trace[:, :, 0].cpu()
            self._host_trace = [(ni, int(tr[ni, 0]), int(tr[ni, 1])) for ni in range(1, sl + 1)]
        return self._host_trace

    @last_trace.setter
    def last_trace(self, v):
        self._host_trace, self._device_state = v, None

    @torch.no_grad()
    def produce_matches(self, data: Dict[str, torch.Tensor], p: float = 0.2, **kwargs):
        d0, d1 = data['descriptors0'], data['descriptors1']
        _lib.require_cuda(d0, 'descriptors0')
        if d0.shape[0] != 1:
            raise ValueError('AdaGML prunes tokens with boolean masks and requires batch size 1 '
                             '(reference nets/adagml.py:358); produce_matches_batched runs a batch of pairs')
        if self._split and self.config.get('device_pruning', True):
            out = self.produce_matches_batched(data, p, check=True)
            return {'matches0': out['matches0'], 'matching_scores0': out['matching_scores0']}
        pk = self.prepare()
        dev = d0.device
        m_full, n_full = d0.shape[1], d1.shape[1]
        cos, sin = self._encode(pk, data)
        ws = self._workspace(m_full + n_full, dev)
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
                    ws = self._workspace(m + n, dev)
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
