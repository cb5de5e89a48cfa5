"""B200-native SegNetViT landmark recogniser -- drop-in for reference ``nets/segnetvit.py``.

``SegNetViT(config).forward(data) -> {'prediction': [B,N,n_class]}`` with the reference's input dict
(``seg_descriptors`` [B,N,256], ``keypoints`` [B,N,2] + ``image`` (shape only) or ``norm_keypoints``),
config keys and state-dict schema (``gnn.layers.i.{qkv,proj,mlp.{0,1,3}}``, ``kenc.Wr``,
``input_proj``, ``seg.{0,1,3}``; reference nets/segnetvit.py:124-203).  The forward pass is a sequence of
libpram_b200 kernel launches; the nn.Module containers only own the parameters.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from .. import _lib, ops
from . import _blocks as B
from .utils import image_wh


class _Layers(nn.Module):
    def __init__(self, n_layers: int, feature_dim: int, hidden_dim: int, num_heads: int):
        super().__init__()
        self.layers = nn.ModuleList([B.SelfBlockParams(feature_dim, hidden_dim, num_heads) for _ in range(n_layers)])


class SegNetViT(nn.Module):
    default_config = {
        'descriptor_dim': 256,
        'output_dim': 1024,
        'n_class': 512,
        'keypoint_encoder': [32, 64, 128, 256],
        'n_layers': 15,
        'num_heads': 4,
        'hidden_dim': 256,
        'with_score': False,
        'with_global': False,
        'with_cls': False,
        'with_sc': False,
    }

    def __init__(self, config: Optional[dict] = None):
        super().__init__()
        self.config = {**self.default_config, **(config or {})}
        c = self.config
        if c['hidden_dim'] != 256 or c['num_heads'] != 4:
            raise _lib.PramError('the sm_100a kernels are specialised for hidden_dim=256, num_heads=4 '
                                 '(the only configuration the reference ships)')
        self.n_layers = c['n_layers']
        self.with_sc = c['with_sc']
        self.gnn = _Layers(c['n_layers'], c['hidden_dim'], c['hidden_dim'], c['num_heads'])
        self.kenc = B.FourierParams(2, c['hidden_dim'] // c['num_heads'])
        self.input_proj = nn.Linear(c['descriptor_dim'], c['hidden_dim'])
        self.seg = B.mlp_holder(c['hidden_dim'], c['output_dim'], c['n_class'])
        if self.with_sc:
            self.sc = B.mlp_holder(c['hidden_dim'], c['output_dim'], 3)
        self._packed = None
        self.precision = 'bf16x3'  # 'bf16x3' | 'bf16' (tcgen05) | 'fp32' (CUDA cores); see nets/sfd2.py
        self.eval()

    def set_precision(self, precision: str, attention_probs: str = 'split'):
        """``attention_probs``: 'split' = the softmax probabilities go to the P.V tensor-core product as bf16 hi / lo planes
        (parity mode), 'f16' = as one IEEE fp16 plane against fp16 hi / lo V planes (csrc/attention_tc.cu, P16: part of the
        mixed mode; tensor-core precisions only)."""
        assert precision in ('bf16x3', 'bf16', 'fp32') and attention_probs in ('split', 'f16')
        self.precision = precision
        self.attention_probs = attention_probs
        return self

    def _workspace(self, tokens: int, device):
        ws = B.Workspace(tokens, device, {'fp32': 0, 'bf16': 1, 'bf16x3': 3}[self.precision])
        ws.p16 = bool(ws.split) and getattr(self, 'attention_probs', 'split') == 'f16' and ops.ATT_P16_ALLOWED
        return ws

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    def prepare(self):
        if self._packed is None:
            if self.input_proj.weight.device.type != 'cuda':
                raise _lib.PramError('SegNetViT must be on a CUDA device: pram_b200 has no CPU path')
            pk = {'layers': [B.pack_self(l) for l in self.gnn.layers],
                  'in.w': B._c(self.input_proj.weight), 'in.b': B._c(self.input_proj.bias),
                  'in.tc': B._tc(self.input_proj.weight),
                  'Wr': B._c(self.kenc.Wr.weight), **B.pack_mlp(self.seg, 'seg')}
            if self.with_sc:
                pk.update(B.pack_mlp(self.sc, 'sc'))
            self._packed = pk
        return self._packed

    @torch.no_grad()
    def forward(self, data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        desc = data['seg_descriptors']
        _lib.require_cuda(desc, 'seg_descriptors')
        pk = self.prepare()
        b, n, dd = desc.shape
        T = b * n
        if 'norm_keypoints' in data:
            cos, sin = ops.posenc(data['norm_keypoints'], 1.0, 1.0, pk['Wr'], prenormalized=True)
        elif 'image' in data:
            w, h = image_wh(data['image'].shape)
            cos, sin = ops.posenc(data['keypoints'], w, h, pk['Wr'])
        else:
            raise ValueError('Require image shape for keypoint coordinate normalization')
        ws = self._workspace(T, desc.device)
        B.input_tokens(ws, pk, desc.reshape(T, dd), 0)
        seg = [(0, b, n)]
        # ``num_keypoints`` [B] (extension for padded batches): tokens >= num_keypoints[b] are padding and get no attention
        cnt = data.get('num_keypoints')
        counts = None
        if cnt is not None:
            counts = [torch.as_tensor(cnt, device=desc.device).to(torch.int32).reshape(-1).contiguous()]
            if counts[0].numel() != b:
                raise ValueError(f'num_keypoints must have one entry per batch element ({b})')
        for lp in pk['layers']:
            B.self_block(ws, lp, seg, cos, sin, counts=counts)
        c = self.config
        out = torch.empty((b, n, c['n_class']), device=desc.device, dtype=torch.float32)
        B.head_mlp(ws, pk, 'seg', c['output_dim'], c['n_class'], out)
        output = {'prediction': out}
        if self.with_sc:
            sc = torch.empty((b, n, 3), device=desc.device, dtype=torch.float32)
            B.head_mlp(ws, pk, 'sc', c['output_dim'], 3, sc)
            output['sc'] = sc
        return output
