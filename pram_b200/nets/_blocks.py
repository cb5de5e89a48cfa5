"""Host-side sequencing of the transformer blocks shared by SegNetViT / GML / AdaGML.

Each function enqueues kernels of libpram_b200 (no torch arithmetic).  Token activations live in the
LEFT half of a [T, 512] "concat" buffer so that ``cat([x, message])`` (reference
nets/segnetvit.py:106, nets/gml.py:137,184) never has to be materialised: the attention projection
writes its message straight into the right half and the MLP reads the 512-wide row.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .. import ops

HEADS = 4
HDIM = 64
D = 256


def mlp_holder(d_in: int, d_hid: int, d_out: int) -> nn.Sequential:
    """Linear / LayerNorm / GELU / Linear parameter holder (indices 0,1,3 as in the reference)."""
    return nn.Sequential(nn.Linear(d_in, d_hid), nn.LayerNorm(d_hid, elementwise_affine=True), nn.GELU(),
                         nn.Linear(d_hid, d_out))


class SelfBlockParams(nn.Module):
    """Parameter holder: reference SelfMultiHeadAttention (nets/segnetvit.py:79-95, nets/gml.py:110-126)."""

    def __init__(self, feat_dim: int = D, hidden_dim: int = D, num_heads: int = HEADS):
        super().__init__()
        self.qkv = nn.Linear(feat_dim, hidden_dim * 3)
        self.proj = nn.Linear(hidden_dim, hidden_dim)
        self.mlp = mlp_holder(feat_dim + hidden_dim, feat_dim * 2, feat_dim)


class CrossBlockParams(nn.Module):
    """Parameter holder: reference CrossMultiHeadAttention (nets/gml.py:143-159)."""

    def __init__(self, feat_dim: int = D, hidden_dim: int = D, num_heads: int = HEADS):
        super().__init__()
        self.to_qk = nn.Linear(feat_dim, hidden_dim)
        self.to_v = nn.Linear(feat_dim, hidden_dim)
        self.proj = nn.Linear(hidden_dim, hidden_dim)
        self.mlp = mlp_holder(feat_dim + hidden_dim, feat_dim * 2, feat_dim)


class FourierParams(nn.Module):
    """Parameter holder: LearnableFourierPositionalEncoding (nets/segnetvit.py:26-33)."""

    def __init__(self, M: int = 2, dim: int = HDIM):
        super().__init__()
        self.Wr = nn.Linear(M, dim // 2, bias=False)
        nn.init.normal_(self.Wr.weight.data, mean=0, std=1.0)


def _c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().float().contiguous()


def pack_self(blk: SelfBlockParams) -> Dict[str, torch.Tensor]:
    """De-interleave the qkv rows once: reference feature index = head*192 + dim*3 + {q,k,v}
    (``unflatten(-1, (heads, -1, 3))``, nets/segnetvit.py:99) -> rows ordered (part, head, dim)."""
    w, b = blk.qkv.weight.detach(), blk.qkv.bias.detach()
    idx = torch.arange(3 * HEADS * HDIM, device=w.device).view(HEADS, HDIM, 3).permute(2, 0, 1).reshape(-1)
    return {'qkv.w': _c(w[idx]), 'qkv.b': _c(b[idx]), 'proj.w': _c(blk.proj.weight), 'proj.b': _c(blk.proj.bias),
            **pack_mlp(blk.mlp, 'mlp')}


def pack_cross(blk: CrossBlockParams) -> Dict[str, torch.Tensor]:
    """to_qk and to_v fused into one [512,256] projection (rows: qk | v)."""
    return {'qkv.w': _c(torch.cat([blk.to_qk.weight.detach(), blk.to_v.weight.detach()], 0)),
            'qkv.b': _c(torch.cat([blk.to_qk.bias.detach(), blk.to_v.bias.detach()], 0)),
            'proj.w': _c(blk.proj.weight), 'proj.b': _c(blk.proj.bias), **pack_mlp(blk.mlp, 'mlp')}


def pack_mlp(mlp: nn.Sequential, pre: str) -> Dict[str, torch.Tensor]:
    return {pre + '.0.w': _c(mlp[0].weight), pre + '.0.b': _c(mlp[0].bias), pre + '.ln.g': _c(mlp[1].weight),
            pre + '.ln.b': _c(mlp[1].bias), pre + '.3.w': _c(mlp[3].weight), pre + '.3.b': _c(mlp[3].bias)}


class Workspace:
    """Per-call scratch for T tokens (all fp32): two concat buffers, qkv, q/k/v, ctx, hidden."""

    def __init__(self, tokens: int, device):
        e = lambda *s: torch.empty(s, device=device, dtype=torch.float32)
        self.T = tokens
        self.cat = [e(tokens, 2 * D), e(tokens, 2 * D)]
        self.cur = 0
        self.qkv = e(tokens, 3 * D)
        self.q, self.k, self.v = e(tokens, D), e(tokens, D), e(tokens, D)
        self.ctx = e(tokens, D)
        self.hid = e(tokens, 2 * D)

    @property
    def x(self) -> torch.Tensor:  # current activations: left half of the current concat buffer
        return self.cat[self.cur]


def run_mlp(pk: Dict[str, torch.Tensor], pre: str, a: torch.Tensor, lda: int, rows: int, d_in: int, d_hid: int,
            d_out: int, hid: torch.Tensor, out: torch.Tensor, ldo: int, res: Optional[torch.Tensor] = None,
            ldres: int = 0) -> torch.Tensor:
    ops.linear_f32(a, lda, pk[pre + '.0.w'], pk[pre + '.0.b'], hid, d_hid, rows, d_in, d_hid)
    ops.layernorm_gelu_(hid, pk[pre + '.ln.g'], pk[pre + '.ln.b'], d_hid)  # hid holds exactly rows x d_hid
    ops.linear_f32(hid, d_hid, pk[pre + '.3.w'], pk[pre + '.3.b'], out, ldo, rows, d_hid, d_out, res=res, ldres=ldres)
    return out


def _finish_block(ws: Workspace, pk: Dict[str, torch.Tensor]):
    """message = proj(ctx) -> right half; x_new = x + mlp([x, message]) -> left half of the other buffer."""
    T = ws.T
    cat = ws.cat[ws.cur]
    nxt = ws.cat[ws.cur ^ 1]
    ops.linear_f32(ws.ctx, D, pk['proj.w'], pk['proj.b'], cat[:, D:], 2 * D, T, D, D)
    run_mlp(pk, 'mlp', cat, 2 * D, T, 2 * D, 2 * D, D, ws.hid, nxt, 2 * D, res=cat, ldres=2 * D)
    ws.cur ^= 1


def self_block(ws: Workspace, pk: Dict[str, torch.Tensor], segments: Sequence[Tuple[int, int, int]],
               cos: torch.Tensor, sin: torch.Tensor, colmeans: Optional[List[torch.Tensor]] = None):
    """One SelfMultiHeadAttention block over all tokens.  ``segments`` = [(token_offset, B, N), ...]:
    attention is computed independently inside each (segment, batch element).
    Reference nets/segnetvit.py:97-106 == nets/gml.py:128-137."""
    T = ws.T
    ops.linear_f32(ws.x, 2 * D, pk['qkv.w'], pk['qkv.b'], ws.qkv, 3 * D, T, D, 3 * D)
    for si, (off, b, n) in enumerate(segments):
        sl = slice(off, off + b * n)
        ops.rotary_split(ws.qkv[sl], 3, b, n, HEADS, cos[sl], sin[sl], 1.0, ws.q[sl], ws.k[sl], ws.v[sl])
        ops.attention_f32(ws.q[sl], ws.k[sl], ws.v[sl], b, HEADS, n, n, HDIM ** -0.5, ws.ctx[sl], D,
                          None if colmeans is None else colmeans[si])
    _finish_block(ws, pk)


def cross_block(ws: Workspace, pk: Dict[str, torch.Tensor], seg0: Tuple[int, int, int], seg1: Tuple[int, int, int],
                colmeans: Optional[List[torch.Tensor]] = None):
    """Bidirectional cross attention between segment 0 (B x M tokens) and segment 1 (B x N tokens) with
    the shared qk projection; both directions reuse the same flash kernel (row softmax of sim and of
    sim^T).  Reference nets/gml.py:164-186.  ``colmeans`` = [mean attn10 over queries -> per token of
    set 0, mean attn01 -> per token of set 1] (reference nets/adagml.py:229)."""
    T = ws.T
    qkv = ws.qkv.view(-1)[:T * 2 * D].view(T, 2 * D)  # (qk | v) rows, 512 wide
    ops.linear_f32(ws.x, 2 * D, pk['qkv.w'], pk['qkv.b'], qkv, 2 * D, T, D, 2 * D)
    (o0, b, m), (o1, _, n) = seg0, seg1
    s0, s1 = slice(o0, o0 + b * m), slice(o1, o1 + b * n)
    sc = (HDIM ** -0.5) ** 0.5  # applied to both qk0 and qk1 (nets/gml.py:174)
    ops.rotary_split(qkv[s0], 2, b, m, HEADS, None, None, sc, ws.q[s0], None, ws.v[s0])
    ops.rotary_split(qkv[s1], 2, b, n, HEADS, None, None, sc, ws.q[s1], None, ws.v[s1])
    # m0 = softmax_rows(sim) v1 ; the column mean of attn01 indexes tokens of set 1
    ops.attention_f32(ws.q[s0], ws.q[s1], ws.v[s1], b, HEADS, m, n, 1.0, ws.ctx[s0], D,
                      None if colmeans is None else colmeans[1])
    # m1 = softmax_rows(sim^T) v0 ; the column mean of attn10 indexes tokens of set 0
    ops.attention_f32(ws.q[s1], ws.q[s0], ws.v[s0], b, HEADS, n, m, 1.0, ws.ctx[s1], D,
                      None if colmeans is None else colmeans[0])
    _finish_block(ws, pk)
