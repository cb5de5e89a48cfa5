"""B200-native GML matcher -- drop-in for reference ``nets/gml.py``.

``GML(config).eval()(data)`` takes the reference's dict (``descriptors0/1`` [B,M|N,128],
``keypoints0/1`` [B,.,2] and one of ``norm_keypoints0/1`` / ``image0/1`` / ``image_shape0/1``) and returns
``matches0/1`` (int64, -1 = unmatched) and ``matching_scores0/1`` (reference nets/gml.py:250-294).
State-dict schema identical to the reference (``bin_score``, ``input_proj``, ``self_attn.i``,
``cross_attn.i``, ``poseenc.Wr``, ``out_proj.i``), so ``weights/imp_gml.920.pth`` loads with
``strict=True``.  The forward pass is kernel launches only: 9 x [self block on both sets, bidirectional
cross block], final projection, M x N distance GEMM, then ONE fused Sinkhorn + match-extraction launch.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from .. import _lib, ops
from . import _blocks as B
from .utils import image_wh


def sinkhorn_matches(dist: torch.Tensor, bin_score: torch.Tensor, iterations: int, p: float = 0.2):
    """dist [B,M,N] -> (matches0, matches1, mscores0, mscores1): fused replacement of the reference's
    ``sink_algorithm`` + ``compute_matches`` (nets/gml.py:27-46, 304-319)."""
    return ops.sinkhorn_match(dist, bin_score, iterations, p)


class GML(nn.Module):
    default_config = {
        'descriptor_dim': 128,
        'hidden_dim': 256,
        'weights': 'indoor',
        'keypoint_encoder': [32, 64, 128, 256],
        'GNN_layers': ['self', 'cross'] * 9,
        'sinkhorn_iterations': 20,
        'match_threshold': 0.2,
        'with_pose': False,
        'n_layers': 9,
        'n_min_tokens': 256,
        'with_sinkhorn': True,
        'ac_fn': 'relu',
        'norm_fn': 'bn',
    }

    def __init__(self, config: Optional[dict] = None):
        super().__init__()
        self.config = {**self.default_config, **(config or {})}
        c = self.config
        if c['hidden_dim'] != 256:
            raise _lib.PramError('the sm_100a kernels are specialised for hidden_dim=256 / 4 heads')
        if not c['with_sinkhorn']:
            raise _lib.PramError('with_sinkhorn=False (dual-softmax scoring) is not on the hot path')
        self.n_layers = c['n_layers']
        self.sinkhorn_iterations = c['sinkhorn_iterations']
        self.match_threshold = c['match_threshold']
        self.input_proj = nn.Linear(c['descriptor_dim'], c['hidden_dim'])
        self.self_attn = nn.ModuleList([B.SelfBlockParams() for _ in range(self.n_layers)])
        self.cross_attn = nn.ModuleList([B.CrossBlockParams() for _ in range(self.n_layers)])
        self.poseenc = B.FourierParams(2, 64)
        self.out_proj = nn.ModuleList([nn.Linear(256, 256) for _ in range(self.n_layers)])
        self.register_parameter('bin_score', nn.Parameter(torch.tensor(1.)))
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

    @property
    def _split(self) -> int:
        return {'fp32': 0, 'bf16': 1, 'bf16x3': 3}[self.precision]

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    def prepare(self):
        if self._packed is None:
            if self.input_proj.weight.device.type != 'cuda':
                raise _lib.PramError('GML must be on a CUDA device: pram_b200 has no CPU path')
            self._packed = {
                'self': [B.pack_self(l) for l in self.self_attn],
                'cross': [B.pack_cross(l) for l in self.cross_attn],
                'in.w': B._c(self.input_proj.weight), 'in.b': B._c(self.input_proj.bias),
                'in.tc': B._tc(self.input_proj.weight),
                'out.tc': [B._tc(l.weight * 0.25) for l in self.out_proj],
                'Wr': B._c(self.poseenc.Wr.weight),
                # out_proj(desc) / d**.25 with d = 256: the factor 1/4 is exact, folded into the weights
                'out': [(B._c(l.weight * 0.25), B._c(l.bias * 0.25)) for l in self.out_proj],
                'bin': self.bin_score.detach().float().reshape(1).contiguous(),
            }
        return self._packed

    def forward(self, data, mode=0):
        if self.training:
            raise _lib.PramError('training is out of scope; call .eval()')
        return self.produce_matches(data)

    # -- helpers shared with AdaGML -----------------------------------------------------------
    def _encode(self, pk, data):
        k0, k1 = data['keypoints0'], data['keypoints1']
        if 'norm_keypoints0' in data and 'norm_keypoints1' in data:
            c0, s0 = ops.posenc(data['norm_keypoints0'], 1., 1., pk['Wr'], True)
            c1, s1 = ops.posenc(data['norm_keypoints1'], 1., 1., pk['Wr'], True)
        elif 'image0' in data and 'image1' in data:
            c0, s0 = ops.posenc(k0, *image_wh(data['image0'].shape), pk['Wr'])
            c1, s1 = ops.posenc(k1, *image_wh(data['image1'].shape), pk['Wr'])
        elif 'image_shape0' in data and 'image_shape1' in data:
            c0, s0 = ops.posenc(k0, *image_wh(data['image_shape0']), pk['Wr'])
            c1, s1 = ops.posenc(k1, *image_wh(data['image_shape1']), pk['Wr'])
        else:
            raise ValueError('Require image shape for keypoint coordinate normalization')
        return torch.cat([c0, c1], 0), torch.cat([s0, s1], 0)

    @staticmethod
    def _counts(c, b: int, device):
        if c is None:
            return None
        c = torch.as_tensor(c, device=device).to(torch.int32).reshape(-1).contiguous()
        if c.numel() != b:
            raise ValueError(f'num_keypoints must have one entry per batch element ({b}), got {c.numel()}')
        return c

    @staticmethod
    def _input_tokens(pk, ws, d0, d1):
        b, m, dd = d0.shape
        n = d1.shape[1]
        B.input_tokens(ws, pk, d0.reshape(b * m, dd), 0)
        B.input_tokens(ws, pk, d1.reshape(b * n, dd), b * m)

    @staticmethod
    def _distance(pk, layer, ws, b, m, n):
        """mdesc = out_proj(desc)/4 for both sets, dist[b] = mdesc0[b] . mdesc1[b]^T (nets/gml.py:278-282)."""
        T = ws.T
        w, bias = pk['out'][layer]
        dist = torch.empty((b, m, n), device=ws.x.device, dtype=torch.float32)
        if ws.split:
            md = ops.empty_split((T, B.D), ws.x.device, ws.split == 3)
            ops.linear_tc(ws.x_bf, 2 * B.D, T, B.D, pk['out.tc'][layer], B.D, bias, out_bf=md, ld_bf=B.D, split=ws.split)
            ops.linear_tc(md, B.D, m, B.D, ops.split_rows(md, b * m), n, out_f32=dist, ld_f32=n, split=ws.split,
                          batch=b, w_batched=True)
        else:
            md = ws.qkv.view(-1)[:T * B.D].view(T, B.D)
            ops.linear_f32(ws.x, 2 * B.D, w, bias, md, B.D, T, B.D, B.D)
            ops.linear_f32(md, B.D, md[b * m:], None, dist, n, m, B.D, n, batch=b, a_bs=m * B.D, w_bs=n * B.D, o_bs=m * n)
        return dist

    @torch.no_grad()
    def produce_matches(self, data: Dict[str, torch.Tensor], p: float = 0.2, **kwargs):
        d0, d1 = data['descriptors0'], data['descriptors1']
        _lib.require_cuda(d0, 'descriptors0')
        pk = self.prepare()
        b, m, _ = d0.shape
        n = d1.shape[1]
        if m == 0 or n == 0:
            raise ValueError('GML needs at least one keypoint per set (the reference fails on empty sets too)')
        cos, sin = self._encode(pk, data)
        ws = self._workspace(b * (m + n), d0.device)
        self._input_tokens(pk, ws, d0, d1)
        seg0, seg1 = (0, b, m), (b * m, b, n)
        # padded batches (extension of the reference dict, which has no batch-with-padding notion: its callers run one pair
        # per call): ``num_keypoints0/1`` [B] = real keypoints of each set; the rest of the [B, M|N] slots are padding that
        # takes no part in attention or in the Sinkhorn normalisation and comes back unmatched
        cnt = [self._counts(data.get('num_keypoints0'), b, d0.device), self._counts(data.get('num_keypoints1'), b, d0.device)]
        counts = cnt if (cnt[0] is not None or cnt[1] is not None) else None
        if counts is not None and m == n:  # [2B] counts of [set 0 | set 1]: both sets go through ONE attention launch per block
            full = lambda c, k: c if c is not None else torch.full((b,), k, device=d0.device, dtype=torch.int32)
            counts = [cnt[0], cnt[1], torch.cat([full(cnt[0], m), full(cnt[1], n)])]
        for i in range(self.n_layers):
            B.self_block(ws, pk['self'][i], (seg0, seg1), cos, sin, counts=counts)
            B.cross_block(ws, pk['cross'][i], seg0, seg1, counts=counts)
        dist = self._distance(pk, self.n_layers - 1, ws, b, m, n)
        m0, m1, s0, s1 = ops.sinkhorn_match(dist, pk['bin'], self.sinkhorn_iterations, p, m_counts=cnt[0], n_counts=cnt[1])
        return {'matches0': m0, 'matches1': m1, 'matching_scores0': s0, 'matching_scores1': s1}
