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

import os

HEADS = 4
HDIM = 64
D = 256
# one fused tcgen05 kernel for proj -> mlp.0 -> LayerNorm + GELU -> mlp.3 (+ residual) (csrc/mlp_block_tc.cu);
# PRAM_FUSED_BLOCK=0 keeps the four separate launches (A/B timing, bisecting)
FUSED_BLOCK = os.environ.get('PRAM_FUSED_BLOCK', '1') != '0'
# PRAM_MERGE_SETS=0: one attention launch per token set / direction as before (A/B timing, bisecting)
MERGE_SETS = os.environ.get('PRAM_MERGE_SETS', '1') != '0'
# AdaGML's per-token mean attention on the tensor cores (attention_tc statistics + pram_attention_colsum_tc);
# PRAM_COLMEAN_TC=0 falls back to the fp32 CUDA-core attention kernel for those blocks (A/B, bisecting)
COLMEAN_TC = os.environ.get('PRAM_COLMEAN_TC', '1') != '0'


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


def _tc(t: torch.Tensor) -> ops.Split:
    """[N,K] Linear weight -> split-bf16 planes (the tensor-core operand layout is torch's own [out,in])."""
    return ops.split_bf16(_c(t))


def pack_self(blk: SelfBlockParams) -> Dict[str, torch.Tensor]:
    """De-interleave the qkv rows once: reference feature index = head*192 + dim*3 + {q,k,v}
    (``unflatten(-1, (heads, -1, 3))``, nets/segnetvit.py:99) -> rows ordered (part, head, dim)."""
    w, b = blk.qkv.weight.detach(), blk.qkv.bias.detach()
    idx = torch.arange(3 * HEADS * HDIM, device=w.device).view(HEADS, HDIM, 3).permute(2, 0, 1).reshape(-1)
    return {'qkv.w': _c(w[idx]), 'qkv.b': _c(b[idx]), 'proj.w': _c(blk.proj.weight), 'proj.b': _c(blk.proj.bias),
            'qkv.tc': _tc(w[idx]), 'proj.tc': _tc(blk.proj.weight), **pack_mlp(blk.mlp, 'mlp'),
            **pack_block_tail(blk.proj, blk.mlp)}


def pack_cross(blk: CrossBlockParams) -> Dict[str, torch.Tensor]:
    """to_qk and to_v fused into one [512,256] projection (rows: qk | v)."""
    wqkv = torch.cat([blk.to_qk.weight.detach(), blk.to_v.weight.detach()], 0)
    return {'qkv.w': _c(wqkv), 'qkv.tc': _tc(wqkv),
            'qkv.b': _c(torch.cat([blk.to_qk.bias.detach(), blk.to_v.bias.detach()], 0)),
            'proj.w': _c(blk.proj.weight), 'proj.b': _c(blk.proj.bias), 'proj.tc': _tc(blk.proj.weight),
            **pack_mlp(blk.mlp, 'mlp'), **pack_block_tail(blk.proj, blk.mlp)}


def pack_block_tail(proj: nn.Linear, mlp: nn.Sequential) -> Dict[str, torch.Tensor]:
    """Operands of the fused block tail: mlp.0([x | proj(ctx)]) = [x | ctx] . W1^T + b1 with
    W1 = [W0[:, :256] | W0[:, 256:] . Wp] and b1 = b0 + W0[:, 256:] . bp, folded once in float64 (like the BatchNorm
    folding of the conv stack); reference nets/segnetvit.py:104-106, nets/gml.py:135-137."""
    if mlp[0].weight.shape != (2 * D, 2 * D) or mlp[3].weight.shape != (D, 2 * D) or proj.weight.shape != (D, D):
        return {}
    w0, b0 = mlp[0].weight.detach().double(), mlp[0].bias.detach().double()
    wp, bp = proj.weight.detach().double(), proj.bias.detach().double()
    w1 = torch.cat([w0[:, :D], w0[:, D:] @ wp], 1)
    b1 = b0 + w0[:, D:] @ bp
    tables = torch.cat([b1.float(), mlp[1].weight.detach().float(), mlp[1].bias.detach().float(),
                        mlp[3].bias.detach().float()]).cpu().contiguous()  # host: becomes a kernel parameter block
    # k-block-major tiles: every TMA box of the kernel is one contiguous run of memory
    w1t = w1.float().view(2 * D, 8, 64).permute(1, 0, 2).contiguous()                           # [8][512][64]
    w3t = mlp[3].weight.detach().float().view(D, 16, 32).permute(1, 0, 2).contiguous()          # [16][256][32]
    return {'blk.w1.tc': ops.split_bf16(w1t), 'blk.w3.tc': ops.split_bf16(w3t), 'blk.tables': tables}


def pack_mlp(mlp: nn.Sequential, pre: str) -> Dict[str, torch.Tensor]:
    d = {pre + '.0.w': _c(mlp[0].weight), pre + '.0.b': _c(mlp[0].bias), pre + '.ln.g': _c(mlp[1].weight),
         pre + '.ln.b': _c(mlp[1].bias), pre + '.3.w': _c(mlp[3].weight), pre + '.3.b': _c(mlp[3].bias)}
    if mlp[0].weight.shape[1] % 8 == 0 and mlp[3].weight.shape[1] % 8 == 0:
        d[pre + '.0.tc'] = _tc(mlp[0].weight)
        d[pre + '.3.tc'] = _tc(mlp[3].weight)
    return d


class Workspace:
    """Per-call scratch for T tokens (all fp32): two concat buffers, qkv, q/k/v, ctx, hidden."""

    def __init__(self, tokens: int, device, split: int = 0):
        e = lambda *s: torch.empty(s, device=device, dtype=torch.float32)
        self.T = tokens
        self.split = split  # 0: fp32 CUDA-core path; 1 / 3: tcgen05 path with bf16 / bf16x3 operands
        self.ctx_in_bf = False
        self.fused = bool(split) and FUSED_BLOCK
        self.keep_f32 = False  # fused path only: also maintain the fp32 activation rows (ws.x) next to the bf16 planes
        # attention probabilities as ONE IEEE fp16 plane against fp16 hi / lo V planes (written so by the qkv epilogue): two
        # PV MMAs per k-step instead of three and about half the softmax instructions, weights exact to 2^-12 relative
        # (set_precision(..., attention_probs='f16')); False: bf16 hi / lo probabilities, bf16 V planes (the parity mode)
        self.p16 = False
        if split:
            lo = split == 3
            self.cat_bf = [ops.empty_split((tokens, 2 * D), device, lo), ops.empty_split((tokens, 2 * D), device, lo)]
            self.ctx_bf = ops.empty_split((tokens, D), device, lo)
            # attention operands written directly by the qkv GEMM epilogue, [B, heads, n, 64] per segment
            self.q_bf = ops.empty_split((tokens, D), device, lo)
            self.k_bf = ops.empty_split((tokens, D), device, lo)
            self.v_bf = ops.empty_split((tokens, D), device, lo)
            self.hid_bf = ops.empty_split((tokens, 2 * D), device, lo)
        self.cat = [e(tokens, 2 * D), e(tokens, 2 * D)]
        self.cur = 0
        self.qkv = e(tokens, 3 * D)
        self.q, self.k, self.v = e(tokens, D), e(tokens, D), e(tokens, D)
        self.ctx = e(tokens, D)
        self.hid = e(tokens, 2 * D)

    def stats(self, b: int, nq: int, nk: int):
        """Scratch of the mean-attention path: (row statistics of the queries [b*heads, lse_ld(nq)], column sums
        [b*heads, lse_ld(nk)]), fresh per call (stream-ordered allocator; a few hundred KB)."""
        dev = self.cat[0].device
        return (torch.empty((b * HEADS, ops.lse_ld(nq)), device=dev, dtype=torch.float32),
                torch.empty((b * HEADS, ops.lse_ld(nk)), device=dev, dtype=torch.float32))

    def ctx_out(self):
        """Where the tensor-core attention writes its context: the right half of the current concat rows when the block
        tail is fused (the projection is folded into mlp.0 there), else the separate [T, 256] buffer."""
        if self.fused:
            return ops.split_cols(self.cat_bf[self.cur], D), 2 * D
        return self.ctx_bf, D

    @property
    def x(self) -> torch.Tensor:  # current activations: left half of the current concat buffer
        return self.cat[self.cur]

    @property
    def x_bf(self) -> ops.Split:
        return self.cat_bf[self.cur]


def linear(ws: Workspace, a_f32, a_bf, lda: int, rows: int, k: int, n: int, pk, name: str, out_f32=None, ld_f32: int = 0,
           out_bf=None, ld_bf: int = 0, res=None, ldres: int = 0):
    """One Linear layer on whichever path the workspace selects (fp32 CUDA cores / tcgen05)."""
    if ws.split:
        ops.linear_tc(a_bf, lda, rows, k, pk[name + '.tc'], n, pk[name + '.b'], res, ldres, False, out_f32, ld_f32,
                      out_bf, ld_bf, split=ws.split)
    else:
        ops.linear_f32(a_f32, lda, pk[name + '.w'], pk[name + '.b'], out_f32, ld_f32, rows, k, n, res=res, ldres=ldres)


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
    if not ws.split:
        ops.linear_f32(ws.ctx, D, pk['proj.w'], pk['proj.b'], cat[:, D:], 2 * D, T, D, D)
        run_mlp(pk, 'mlp', cat, 2 * D, T, 2 * D, 2 * D, D, ws.hid, nxt, 2 * D, res=cat, ldres=2 * D)
    else:
        cbf, nbf = ws.cat_bf[ws.cur], ws.cat_bf[ws.cur ^ 1]
        if not ws.ctx_in_bf:  # CUDA-core attention produced fp32 ctx (AdaGML's mean-attention variant)
            ops.split_bf16_into(ws.ctx, ws.ctx_bf)
        if ws.fused:
            if not ws.ctx_in_bf:  # layout plumbing: the fused kernel reads ctx from the right half of the concat rows
                cbf.hi[:, D:].copy_(ws.ctx_bf.hi)
                if cbf.lo is not None:
                    cbf.lo[:, D:].copy_(ws.ctx_bf.lo)
            # the fp32 copy of the activations is only kept when a CUDA-core consumer needs it (AdaGML's pooling / compaction);
            # otherwise the stream lives as its two bf16 planes and the kernel takes the residual from them
            f32 = ws.keep_f32
            ops.mlp_block_tc(cbf, 2 * D, T, pk['blk.w1.tc'], pk['blk.w3.tc'], pk['blk.tables'], cat if f32 else None, 2 * D,
                             nxt if f32 else None, 2 * D, nbf, 2 * D, split=ws.split)
            ws.cur ^= 1
            return
        linear(ws, None, ws.ctx_bf, D, T, D, D, pk, 'proj', out_bf=ops.split_cols(cbf, D), ld_bf=2 * D)
        linear(ws, None, cbf, 2 * D, T, 2 * D, 2 * D, pk, 'mlp.0', out_f32=ws.hid, ld_f32=2 * D)
        ops.layernorm_gelu_split(ws.hid, pk['mlp.ln.g'], pk['mlp.ln.b'], 2 * D, ws.hid_bf)
        linear(ws, None, ws.hid_bf, 2 * D, T, 2 * D, D, pk, 'mlp.3', out_f32=nxt, ld_f32=2 * D, out_bf=nbf, ld_bf=2 * D,
               res=cat, ldres=2 * D)
    ws.cur ^= 1


def _no_counts_here(counts):
    if counts is not None and any(c is not None for c in counts):
        from .. import _lib
        raise _lib.PramError('per-frame keypoint counts (padded batches) are only supported on the tensor-core attention path')


def _mergeable(segments, colmeans, counts) -> bool:
    """Both token sets in ONE attention launch: two adjacent segments of the same shape, no mean-attention output, and
    (for padded batches) the concatenated counts supplied as ``counts[2]`` ([2 B] int32 = counts of set 0, then of set 1)."""
    if not MERGE_SETS or colmeans is not None or len(segments) != 2:
        return False
    (o0, b0, n0), (o1, b1, n1) = segments
    if b0 != b1 or n0 != n1 or o1 != o0 + b0 * n0:
        return False
    return counts is None or (len(counts) > 2 and counts[2] is not None)


def self_block(ws: Workspace, pk: Dict[str, torch.Tensor], segments: Sequence[Tuple[int, int, int]],
               cos: torch.Tensor, sin: torch.Tensor, colmeans: Optional[List[torch.Tensor]] = None,
               counts: Optional[Sequence[Optional[torch.Tensor]]] = None):
    """One SelfMultiHeadAttention block over all tokens.  ``segments`` = [(token_offset, B, N), ...]:
    attention is computed independently inside each (segment, batch element).  ``counts[s]`` (optional, [B] int32 per
    segment): tokens >= counts[s][b] of batch element b are padding and receive no attention.
    Reference nets/segnetvit.py:97-106 == nets/gml.py:128-137."""
    T = ws.T
    use_tc = bool(ws.split) and (colmeans is None or COLMEAN_TC)
    if not use_tc:
        _no_counts_here(counts)
    ws.ctx_in_bf = use_tc
    if use_tc:
        # qkv GEMM with the fused epilogue: bias, rotary on q/k, split-bf16 Q/K/V in [B, heads, n, 64]
        seg_split = segments[1][0] if len(segments) > 1 else T
        qkv = {'mode': 1, 'scale': 1.0, 'cos': cos, 'sin': sin, 'q': ws.q_bf, 'k': ws.k_bf, 'v': ws.v_bf,
               'seg_split': seg_split, 'seg_n0': segments[0][2], 'seg_n1': segments[-1][2], 'v_f16': ws.p16}
        ops.linear_tc(ws.x_bf, 2 * D, T, D, pk['qkv.tc'], 3 * D, pk['qkv.b'], split=ws.split, bn=256, qkv=qkv)
        ctx, ctx_ld = ws.ctx_out()
        if _mergeable(segments, colmeans, counts):
            # two equally shaped, adjacent segments are one batch of 2 B elements: one launch (2048 work items at the bench
            # shape fill the 296 CTA slots 6.9 times; two launches of 1024 are 3.5 waves each, rounded up to 4)
            off, b, n = segments[0]
            ops.attention_tc(ops.split_rows(ws.q_bf, off), ops.split_rows(ws.k_bf, off), ops.split_rows(ws.v_bf, off), 2 * b, HEADS,
                             n, n, n, HDIM ** -0.5, None, ops.split_rows(ctx, off), ctx_ld, ws.split, v_mn=True, v_f16=ws.p16,
                             nk_counts=None if counts is None else counts[2])
            _finish_block(ws, pk)
            return
        for si, (off, b, n) in enumerate(segments):
            q, k = ops.split_rows(ws.q_bf, off), ops.split_rows(ws.k_bf, off)
            cnt = None if counts is None else counts[si]
            lse, colsum = ws.stats(b, n, n) if colmeans is not None else (None, None)
            ops.attention_tc(q, k, ops.split_rows(ws.v_bf, off), b, HEADS, n, n, n, HDIM ** -0.5, None,
                             ops.split_rows(ctx, off), ctx_ld, ws.split, v_mn=True, v_f16=ws.p16, nk_counts=cnt, lse_out=lse)
            if colmeans is not None:  # AdaGML: mean attention per key token from S^T tiles + the queries' row statistics
                ops.attention_colmean_tc(k, q, b, HEADS, n, n, HDIM ** -0.5, lse, colsum, colmeans[si], colmeans[si].stride(0),
                                         ws.split, nq_counts=cnt)
        _finish_block(ws, pk)
        return
    linear(ws, ws.x, ws.x_bf if ws.split else None, 2 * D, T, D, 3 * D, pk, 'qkv', out_f32=ws.qkv, ld_f32=3 * D)
    for si, (off, b, n) in enumerate(segments):
        sl = slice(off, off + b * n)
        ops.rotary_split(ws.qkv[sl], 3, b, n, HEADS, cos[sl], sin[sl], 1.0, ws.q[sl], ws.k[sl], ws.v[sl])
        ops.attention_f32(ws.q[sl], ws.k[sl], ws.v[sl], b, HEADS, n, n, HDIM ** -0.5, ws.ctx[sl], D,
                          None if colmeans is None else colmeans[si])
    _finish_block(ws, pk)


def cross_block(ws: Workspace, pk: Dict[str, torch.Tensor], seg0: Tuple[int, int, int], seg1: Tuple[int, int, int],
                colmeans: Optional[List[torch.Tensor]] = None, counts: Optional[Sequence[Optional[torch.Tensor]]] = None):
    """Bidirectional cross attention between segment 0 (B x M tokens) and segment 1 (B x N tokens) with
    the shared qk projection; both directions reuse the same flash kernel (row softmax of sim and of
    sim^T).  Reference nets/gml.py:164-186.  ``colmeans`` = [mean attn10 over queries -> per token of
    set 0, mean attn01 -> per token of set 1] (reference nets/adagml.py:229)."""
    T = ws.T
    (o0, b, m), (o1, _, n) = seg0, seg1
    s0, s1 = slice(o0, o0 + b * m), slice(o1, o1 + b * n)
    sc = (HDIM ** -0.5) ** 0.5  # applied to both qk0 and qk1 (nets/gml.py:174)
    use_tc = bool(ws.split) and (colmeans is None or COLMEAN_TC)
    ws.ctx_in_bf = use_tc
    if not use_tc:
        _no_counts_here(counts)
    c0, c1 = (None, None) if counts is None else (counts[0], counts[1])
    if use_tc:
        fused = {'mode': 2, 'scale': sc, 'q': ws.q_bf, 'v': ws.v_bf, 'seg_split': o1, 'seg_n0': m, 'seg_n1': n, 'v_f16': ws.p16}
        ops.linear_tc(ws.x_bf, 2 * D, T, D, pk['qkv.tc'], 2 * D, pk['qkv.b'], split=ws.split, bn=256, qkv=fused)
        q0, q1 = ops.split_rows(ws.q_bf, o0), ops.split_rows(ws.q_bf, o1)
        v0, v1 = ops.split_rows(ws.v_bf, o0), ops.split_rows(ws.v_bf, o1)
        ctx, ctx_ld = ws.ctx_out()
        if _mergeable((seg0, seg1), colmeans, counts):
            # both directions in one launch over [set 0 | set 1]: query element i attends to the keys / values (and key
            # counts) of element (i + B) mod 2B, i.e. set 0 -> set 1 and set 1 -> set 0 (pram_attention_tc_shift)
            ops.attention_tc(q0, q0, v0, 2 * b, HEADS, m, n, n, 1.0, None, ops.split_rows(ctx, o0), ctx_ld, ws.split, v_mn=True,
                             v_f16=ws.p16, nk_counts=None if counts is None else counts[2], kv_shift=b)
            _finish_block(ws, pk)
            return
        want = colmeans is not None
        lse, colsum = ws.stats(b, m, n) if want else (None, None)
        ops.attention_tc(q0, q1, v1, b, HEADS, m, n, n, 1.0, None, ops.split_rows(ctx, o0), ctx_ld, ws.split, v_mn=True, v_f16=ws.p16,
                         nk_counts=c1, lse_out=lse)
        if want:  # mean of attn01 over the queries of set 0 -> per token of set 1
            ops.attention_colmean_tc(q1, q0, b, HEADS, n, m, 1.0, lse, colsum, colmeans[1], colmeans[1].stride(0), ws.split,
                                     nq_counts=c0)
            lse, colsum = ws.stats(b, n, m)
        ops.attention_tc(q1, q0, v0, b, HEADS, n, m, m, 1.0, None, ops.split_rows(ctx, o1), ctx_ld, ws.split, v_mn=True, v_f16=ws.p16,
                         nk_counts=c0, lse_out=lse)
        if want:  # mean of attn10 over the queries of set 1 -> per token of set 0
            ops.attention_colmean_tc(q0, q1, b, HEADS, m, n, 1.0, lse, colsum, colmeans[0], colmeans[0].stride(0), ws.split,
                                     nq_counts=c1)
        _finish_block(ws, pk)
        return
    qkv = ws.qkv.view(-1)[:T * 2 * D].view(T, 2 * D)  # (qk | v) rows, 512 wide
    linear(ws, ws.x, ws.x_bf if ws.split else None, 2 * D, T, D, 2 * D, pk, 'qkv', out_f32=qkv, ld_f32=2 * D)
    ops.rotary_split(qkv[s0], 2, b, m, HEADS, None, None, sc, ws.q[s0], None, ws.v[s0])
    ops.rotary_split(qkv[s1], 2, b, n, HEADS, None, None, sc, ws.q[s1], None, ws.v[s1])
    # m0 = softmax_rows(sim) v1 ; the column mean of attn01 indexes tokens of set 1
    ops.attention_f32(ws.q[s0], ws.q[s1], ws.v[s1], b, HEADS, m, n, 1.0, ws.ctx[s0], D,
                      None if colmeans is None else colmeans[1])
    # m1 = softmax_rows(sim^T) v0 ; the column mean of attn10 indexes tokens of set 0
    ops.attention_f32(ws.q[s1], ws.q[s0], ws.v[s0], b, HEADS, n, m, 1.0, ws.ctx[s1], D,
                      None if colmeans is None else colmeans[0])
    _finish_block(ws, pk)


def input_tokens(ws: Workspace, pk, x: torch.Tensor, row0: int):
    """input_proj: x [rows, dd] fp32 -> rows [row0, row0+rows) of the current activation buffers."""
    rows, dd = x.shape
    x = x.float()
    x = x if x.is_contiguous() else x.contiguous()
    if ws.split:
        ops.linear_tc(ops.split_bf16(x, ws.split == 3), dd, rows, dd, pk['in.tc'], D, pk['in.b'], None, 0, False,
                      ws.x[row0:], 2 * D, ops.split_rows(ws.x_bf, row0), 2 * D, split=ws.split)
    else:
        ops.linear_f32(x, dd, pk['in.w'], pk['in.b'], ws.x[row0:], 2 * D, rows, dd, D)


def head_mlp(ws: Workspace, pk, pre: str, d_hid: int, d_out: int, out: torch.Tensor):
    """Linear(256->d_hid) -> LN -> GELU -> Linear(d_hid->d_out) on the current activations (seg head)."""
    T = ws.T
    hid = torch.empty((T, d_hid), device=out.device, dtype=torch.float32)
    if ws.split and (pre + '.0.tc') in pk:
        linear(ws, None, ws.x_bf, 2 * D, T, D, d_hid, pk, pre + '.0', out_f32=hid, ld_f32=d_hid)
        hb = ops.empty_split((T, d_hid), out.device, ws.split == 3)
        ops.layernorm_gelu_split(hid, pk[pre + '.ln.g'], pk[pre + '.ln.b'], d_hid, hb)
        linear(ws, None, hb, d_hid, T, d_hid, d_out, pk, pre + '.3', out_f32=out, ld_f32=d_out)
    else:
        run_mlp(pk, pre, ws.x, 2 * D, T, D, d_hid, d_out, hid, out, d_out)
    return out
