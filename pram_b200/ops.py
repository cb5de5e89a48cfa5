"""Thin tensor-level wrappers over the C ABI (one function per kernel group of SURVEY.md section 2b).

Everything here takes CUDA tensors, allocates outputs through torch, and launches on torch's
current stream.  No arithmetic happens in Python.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream_ptr

Tensor = torch.Tensor


def _f32c(t: Tensor) -> Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


# ---- K5 / K5b ---------------------------------------------------------------------------------

def score_map(logits_nhwc: Tensor, ih: Optional[int] = None, iw: Optional[int] = None) -> Tensor:
    """logits [B,Hc,Wc,>=65] (any strides, channel-last view allowed) -> score [B,Hc*8,Wc*8]
    (optionally bilinearly resized to [B,ih,iw]); reference nets/sfd2.py:294-303."""
    _lib.require_cuda(logits_nhwc, 'logits')
    b, hc, wc, _ = logits_nhwc.shape
    sb, sy, sx, sc = logits_nhwc.stride()
    out = torch.empty((b, hc * 8, wc * 8), device=logits_nhwc.device, dtype=torch.float32)
    call('pram_score_map', ptr(logits_nhwc), sb, sy, sx, sc, b, hc, wc, ptr(out), stream_ptr())
    if ih is not None and (ih != hc * 8 or iw != wc * 8):
        res = torch.empty((b, ih, iw), device=out.device, dtype=torch.float32)
        call('pram_resize_bilinear', ptr(out), b, hc * 8, wc * 8, ptr(res), ih, iw, stream_ptr())
        out = res
    return out


# ---- K6 / K7 ----------------------------------------------------------------------------------

def default_cand_cap(h: int, w: int, radius: int = 4) -> int:
    """Candidate-buffer size per frame: local maxima are >= radius + 1 apart unless the map has exact plateaus."""
    return max(4096, ((h + radius) // (radius + 1)) * ((w + radius) // (radius + 1)))


def detect_keypoints(score: Tensor, conf_th: float, min_keypoints: int, max_keypoints: int, border: int,
                     radius: int = 4, strict: bool = False, fallback: bool = True,
                     return_nms: bool = False, cap: Optional[int] = None, border_hi: Optional[Tuple[int, int]] = None,
                     return_valid: bool = False):
    """score [B,H,W] -> (kpts [B,kpad,2] (x,y) f32, scores [B,kpad], n [B] int32, cand_count [B] [, nms [B,H,W]]
    [, n_valid [B]]).  ``cand_count`` > the candidate cap (``default_cand_cap`` unless ``cap`` is given) means NMS
    survivors were dropped (plateau image) and the call must be repeated with ``cap = H * W``; ``n_valid`` > kpad means
    "unlimited" selection (max_keypoints < 0) found more keypoints than the kernel's 4096 slots and returned the best 4096.
    ``border_hi`` = exclusive upper bounds (y_hi, x_hi) of the border window when they are not (H - border, W - border).

    NMS + threshold + border + top-k on the device with no host sync (reference nets/sfd2.py:305-329).
    ``strict`` selects the ``>`` comparison of the export path (nets/sfd2.py:435); ``fallback`` the
    halve-the-threshold rule of nets/sfd2.py:311-315 (per frame).
    """
    _lib.require_cuda(score, 'score map')
    score = _f32c(score)
    b, h, w = score.shape
    th_hi = np.float32(conf_th)
    if strict:  # s > th  <=>  s >= nextafter(th)
        th_hi = np.nextafter(th_hi, np.float32(np.inf))
    th_lo = np.float32(conf_th * 0.5) if fallback else th_hi
    if strict and fallback:
        th_lo = np.nextafter(th_lo, np.float32(np.inf))
    kmax = 4096
    if max_keypoints > kmax:
        raise _lib.PramError(f'max_keypoints > {kmax} is not supported by the selection kernel')
    kpad = max_keypoints if max_keypoints >= 0 else kmax
    kpad = max(kpad, 1)
    if cap is None:
        cap = default_cand_cap(h, w, radius)
    dev = score.device
    cand = torch.empty((b, cap), device=dev, dtype=torch.int64)
    counts = torch.empty((2, b), device=dev, dtype=torch.int32)
    nms = torch.empty_like(score) if return_nms else None
    call('pram_nms_candidates', ptr(score), b, h, w, radius, float(th_lo), float(th_hi), ptr(nms), ptr(cand), cap,
         ptr(counts[0]), ptr(counts[1]), stream_ptr())
    kpts = torch.empty((b, kpad, 2), device=dev, dtype=torch.float32)
    scs = torch.empty((b, kpad), device=dev, dtype=torch.float32)
    n = torch.empty((b,), device=dev, dtype=torch.int32)
    n_valid = torch.empty((b,), device=dev, dtype=torch.int32) if return_valid else None
    y_hi, x_hi = border_hi if border_hi is not None else (0, 0)
    call('pram_select_keypoints', ptr(cand), cap, ptr(counts[0]), ptr(counts[1]), ptr(score), b, h, w,
         float(th_lo), float(th_hi), int(min_keypoints) if fallback else -1, int(max_keypoints), int(border),
         int(y_hi), int(x_hi), ptr(kpts), ptr(scs), ptr(n), kpad, ptr(n_valid), stream_ptr())
    out = (kpts, scs, n, counts[0])
    if return_nms:
        out = out + (nms,)
    if return_valid:
        out = out + (n_valid,)
    return out


# ---- K8 ---------------------------------------------------------------------------------------

def sample_features(fmap_nhwc: Tensor, kpts: Tensor, counts: Optional[Tensor], s: int, normalize: bool) -> Tensor:
    """fmap [B,h,w,C] contiguous NHWC, kpts [B,kpad,2] -> [B,kpad,C]; reference nets/sfd2.py:53-64."""
    _lib.require_cuda(fmap_nhwc, 'feature map')
    assert fmap_nhwc.is_contiguous() and fmap_nhwc.dtype == torch.float32
    b, h, w, c = fmap_nhwc.shape
    kpts = _f32c(kpts)
    kpad = kpts.shape[1]
    out = torch.empty((b, kpad, c), device=fmap_nhwc.device, dtype=torch.float32)
    call('pram_sample_features', ptr(fmap_nhwc), b, c, h, w, ptr(kpts), ptr(counts), kpad, int(s),
         int(bool(normalize)), ptr(out), stream_ptr())
    return out


def gather_scores(score: Tensor, kpts: Tensor, counts: Optional[Tensor]) -> Tensor:
    score = _f32c(score)
    b, h, w = score.shape
    kpts = _f32c(kpts)
    out = torch.empty(kpts.shape[:2], device=score.device, dtype=torch.float32)
    call('pram_gather_scores', ptr(score), b, h, w, ptr(kpts), ptr(counts), kpts.shape[1], ptr(out), stream_ptr())
    return out


# ---- K9 ---------------------------------------------------------------------------------------

def posenc(kpts: Tensor, width: float, height: float, wr: Tensor, prenormalized: bool = False) -> Tuple[Tensor, Tensor]:
    """kpts [..., 2] -> (cos, sin) each [tokens, 32]; reference nets/utils.py:17-24 + segnetvit.py:35-40."""
    kpts = _f32c(kpts)
    tokens = kpts.numel() // 2
    cos = torch.empty((tokens, 32), device=kpts.device, dtype=torch.float32)
    sin = torch.empty_like(cos)
    call('pram_posenc', ptr(kpts), tokens, float(width), float(height), int(prenormalized), ptr(wr), ptr(cos),
         ptr(sin), stream_ptr())
    return cos, sin


# ---- fp32 network building blocks ----------------------------------------------------------------

def conv_f32(x_nhwc: Tensor, w: Tensor, bias: Optional[Tensor], ksize: int, stride: int, relu: bool,
             res: Optional[Tensor] = None, out: Optional[Tensor] = None) -> Tensor:
    """x [B,H,W,Cin] NHWC contiguous, w [taps,Cin,Cout] -> [B,Ho,Wo,Cout]."""
    b, h, wd, cin = x_nhwc.shape
    cout = w.shape[-1]
    pad = ksize // 2
    ho = (h + 2 * pad - ksize) // stride + 1
    wo = (wd + 2 * pad - ksize) // stride + 1
    if out is None:
        out = torch.empty((b, ho, wo, cout), device=x_nhwc.device, dtype=torch.float32)
    call('pram_conv_f32', ptr(x_nhwc), cin, ptr(w), ptr(bias), ptr(res), cout, ptr(out), cout, b, h, wd, cin, cout,
         ksize, stride, int(relu), stream_ptr())
    return out


def gconv3x3_f32(x_nhwc: Tensor, w: Tensor, bias: Tensor, relu: bool) -> Tensor:
    b, h, wd, c = x_nhwc.shape
    out = torch.empty_like(x_nhwc)
    call('pram_gconv3x3_f32', ptr(x_nhwc), ptr(w), ptr(bias), ptr(out), b, h, wd, c // 8, int(relu), stream_ptr())
    return out


def linear_f32(a: Tensor, lda: int, w: Tensor, bias: Optional[Tensor], out: Tensor, ldo: int, rows: int, k: int,
               n: int, relu: bool = False, res: Optional[Tensor] = None, ldres: int = 0, batch: int = 1,
               a_bs: int = 0, w_bs: int = 0, o_bs: int = 0) -> Tensor:
    """out[rows, n] = a[rows, k] @ w[n, k]^T (+bias)(+res)(relu); raw row strides so that inputs and
    outputs can be column slices of wider buffers (concat-free MLP)."""
    call('pram_linear_f32', ptr(a), lda, ptr(w), ptr(bias), ptr(res), ldres, ptr(out), ldo, rows, k, n, int(relu),
         batch, a_bs, w_bs, o_bs, stream_ptr())
    return out


def l2norm_rows_(x: Tensor, c: int) -> Tensor:
    rows = x.numel() // c
    call('pram_l2norm_rows', ptr(x), ptr(x), rows, c, stream_ptr())
    return x


def layernorm_gelu_(x: Tensor, gamma: Tensor, beta: Tensor, c: int, gelu: bool = True) -> Tensor:
    rows = x.numel() // c
    call('pram_layernorm_gelu', ptr(x), ptr(gamma), ptr(beta), ptr(x), rows, c, int(gelu), stream_ptr())
    return x


def rotary_split(qkv: Tensor, nparts: int, b: int, n: int, heads: int, cos: Optional[Tensor], sin: Optional[Tensor],
                 scale_qk: float, q: Tensor, k: Optional[Tensor], v: Tensor):
    call('pram_rotary_split', ptr(qkv), nparts, b, n, heads, ptr(cos), ptr(sin), float(scale_qk), ptr(q), ptr(k),
         ptr(v), stream_ptr())


def attention_f32(q: Tensor, k: Tensor, v: Tensor, b: int, heads: int, nq: int, nk: int, scale: float, out: Tensor,
                  out_stride: int, colmean: Optional[Tensor] = None):
    ws = None
    if colmean is not None:
        ws = torch.empty(int(_lib.load().pram_attention_f32_colmean_ws_floats(b, heads, nq, nk)), device=q.device, dtype=torch.float32)
    call('pram_attention_f32', ptr(q), ptr(k), ptr(v), b, heads, nq, nk, float(scale), ptr(out), out_stride,
         ptr(colmean), ptr(ws), stream_ptr())


# ---- K15 / K16 --------------------------------------------------------------------------------

def sinkhorn_match(dist: Tensor, bin_score: Tensor, iters: int, threshold: float, cluster: int = 0,
                   return_P: bool = False, m_counts: Optional[Tensor] = None, n_counts: Optional[Tensor] = None):
    """dist [B,M,N] f32, bin_score 0-dim device tensor -> matches0 [B,M] i64, matches1 [B,N] i64,
    mscores0 [B,M], mscores1 [B,N] (+ P [B,M+1,N+1]); reference nets/gml.py:27-46, 304-319.
    ``m_counts`` / ``n_counts`` [B] int32: pair b is the problem of its first m_counts[b] rows / n_counts[b] columns
    (plus the dustbins); the rest of the block is padding."""
    _lib.require_cuda(dist, 'dist')
    dist = _f32c(dist)
    b, m, n = dist.shape
    dev = dist.device
    ldp = (n + 1 + 3) // 4 * 4
    pws = torch.empty((b, m + 1, ldp), device=dev, dtype=torch.float32)
    iws = torch.empty((b * (m + n),), device=dev, dtype=torch.int32)
    fws = torch.empty((b * m,), device=dev, dtype=torch.float32)
    m0 = torch.empty((b, m), device=dev, dtype=torch.int64)
    m1 = torch.empty((b, n), device=dev, dtype=torch.int64)
    s0 = torch.empty((b, m), device=dev, dtype=torch.float32)
    s1 = torch.empty((b, n), device=dev, dtype=torch.float32)
    bs = bin_score.detach().reshape(1).float().contiguous()
    call('pram_sinkhorn_match', ptr(dist), b, m, n, ptr(bs), int(iters), float(threshold), ptr(pws), ptr(iws),
         ptr(fws), ptr(m0), ptr(m1), ptr(s0), ptr(s1), int(cluster), ptr(m_counts), ptr(n_counts), stream_ptr())
    if return_P:
        return m0, m1, s0, s1, pws[:, :, :n + 1]
    return m0, m1, s0, s1


# ---- tensor-core (tcgen05) implicit GEMM ----------------------------------------------------------

class Split:
    """An activation / weight tensor as error-compensated bf16 planes: x ~= hi + lo."""
    __slots__ = ('hi', 'lo')

    def __init__(self, hi: Tensor, lo: Optional[Tensor]):
        self.hi, self.lo = hi, lo

    @property
    def shape(self):
        return self.hi.shape

    def float(self) -> Tensor:
        return self.hi.float() + (self.lo.float() if self.lo is not None else 0)


def as_f16_plane(x: Tensor) -> Split:
    """fp32 -> ONE plane of IEEE fp16 bits (held in a bfloat16-typed tensor: the TMA descriptors move 2-byte elements and
    do not care) -- operand of the single-pass fp16 mode (``f16=True``, ``split=1``)."""
    x = _f32c(x)
    out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    call('pram_cast_f16', ptr(x), ptr(out), x.numel(), stream_ptr())
    return Split(out, None)


def split_bf16(x: Tensor, with_lo: bool = True) -> Split:
    x = _f32c(x)
    hi = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    lo = torch.empty_like(hi) if with_lo else None
    call('pram_split_bf16', ptr(x), ptr(hi), ptr(lo), x.numel(), stream_ptr())
    return Split(hi, lo)


def split_f16(x: Tensor, with_lo: bool = True) -> Split:
    """fp32 -> IEEE fp16 hi / lo planes (held in bfloat16-typed tensors: 2-byte elements for the TMA descriptors) -- the V
    operand of ``attention_tc(v_f16=True)``."""
    x = _f32c(x)
    hi = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    lo = torch.empty_like(hi) if with_lo else None
    call('pram_split_f16', ptr(x), ptr(hi), ptr(lo), x.numel(), stream_ptr())
    return Split(hi, lo)


def split_bf16_into(x: Tensor, out: Split) -> Split:
    """Same as split_bf16 but into preallocated planes (x contiguous fp32, same numel)."""
    call('pram_split_bf16', ptr(x), ptr(out.hi), ptr(out.lo), x.numel(), stream_ptr())
    return out


def split_cols(s: Split, start: int) -> Split:
    """Column-offset view of 2-D split planes (keeps the parent's row stride)."""
    return Split(s.hi[:, start:], s.lo[:, start:] if s.lo is not None else None)


def split_rows(s: Split, start: int) -> Split:
    return Split(s.hi[start:], s.lo[start:] if s.lo is not None else None)


def empty_split(shape, device, with_lo: bool = True, zero: bool = False) -> Split:
    mk = torch.zeros if zero else torch.empty
    hi = mk(shape, device=device, dtype=torch.bfloat16)
    return Split(hi, mk(shape, device=device, dtype=torch.bfloat16) if with_lo else None)


import os as _os
GEMM_CLUSTER = int(_os.environ.get('PRAM_GEMM_CLUSTER', '0'))  # 0 = auto; 1 / 2 force single CTAs / 2-CTA clusters (tests, A/B timing)
GEMM_L2_PREFETCH = int(_os.environ.get('PRAM_GEMM_L2_PREFETCH', '0'))  # 1 = on (experiment; default off)
_S1_TAPS = [(dy, dx, 0) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
# stride 2 on a 2x2 phase-split input: tap r in {0,1,2} reads phase (1,0,1) at offset (-1,0,0)
_PH = ((1, -1), (0, 0), (1, 0))
_S2_TAPS = [(_PH[r][1], _PH[s][1], _PH[r][0] * 2 + _PH[s][0]) for r in range(3) for s in range(3)]


def gemm_tc(a: Split, a_ld: int, in_w: int, in_h: int, in_planes: int, cin: int, w: Split, w_planes: int,
            b: int, ho: int, wo: int, n: int, taps, planes_per_image: int = 1, tw_log2: int = 7,
            w_batch_mult: int = 0, bias: Optional[Tensor] = None, res: Optional[Tensor] = None, res_ld: int = 0,
            relu: bool = False, out_f32: Optional[Tensor] = None, ld_f32: int = 0, out_bf: Optional[Split] = None,
            ld_bf: int = 0, out_ps: Optional[Split] = None, ld_ps: int = 0, l2norm: bool = False, split: int = 3,
            bn: int = 0, qkv: Optional[dict] = None, f16: bool = False, res_bf: Optional[Split] = None,
            out_h16: Optional[Tensor] = None):
    """Raw launch of the tcgen05 implicit-GEMM kernel (see include/pram_b200.h: pram_gemm_tc)."""
    A = _lib.TcArgs()
    A.a_hi, A.a_lo, A.a_ld = a.hi.data_ptr(), (a.lo.data_ptr() if a.lo is not None else None), a_ld
    A.in_W, A.in_H, A.in_planes, A.Cin = in_w, in_h, in_planes, cin
    A.w_hi, A.w_lo, A.w_planes = w.hi.data_ptr(), (w.lo.data_ptr() if w.lo is not None else None), w_planes
    A.B, A.Ho, A.Wo, A.N = b, ho, wo, n
    A.tw_log2, A.ntaps = tw_log2, len(taps)
    for i, (dy, dx, pl) in enumerate(taps):
        A.tap_dy[i], A.tap_dx[i], A.tap_plane[i] = dy, dx, pl
    A.planes_per_image, A.w_batch_mult = planes_per_image, w_batch_mult
    A.bias = bias.data_ptr() if bias is not None else None
    A.res, A.res_ld = (res.data_ptr() if res is not None else None), res_ld
    A.relu = int(relu)
    A.out_f32, A.ld_f32 = (out_f32.data_ptr() if out_f32 is not None else None), ld_f32
    if out_bf is not None:
        A.out_hi, A.out_lo, A.ld_bf = out_bf.hi.data_ptr(), (out_bf.lo.data_ptr() if out_bf.lo is not None else None), ld_bf
    if out_ps is not None:
        A.ps_hi, A.ps_lo, A.ld_ps = out_ps.hi.data_ptr(), (out_ps.lo.data_ptr() if out_ps.lo is not None else None), ld_ps
    A.l2norm, A.split, A.bn = int(l2norm), split, bn
    if qkv is not None:
        A.qkv_mode, A.qk_scale, A.heads = qkv['mode'], float(qkv['scale']), qkv.get('heads', 4)
        A.cosb = qkv['cos'].data_ptr() if qkv.get('cos') is not None else None
        A.sinb = qkv['sin'].data_ptr() if qkv.get('sin') is not None else None
        for name in ('q', 'k', 'v'):
            sp = qkv.get(name)
            if sp is not None:
                setattr(A, name + '_hi', sp.hi.data_ptr())
                setattr(A, name + '_lo', sp.lo.data_ptr() if sp.lo is not None else None)
        A.seg_split, A.seg_n0, A.seg_n1 = qkv['seg_split'], qkv['seg_n0'], qkv['seg_n1']
        A.v_f16 = int(bool(qkv.get('v_f16', False)))
    A.cluster = GEMM_CLUSTER
    A.l2_prefetch = GEMM_L2_PREFETCH
    A.f16 = int(f16)
    if res_bf is not None:  # residual from split-bf16 planes instead of an fp32 tensor
        A.res_hi, A.res_lo = res_bf.hi.data_ptr(), (res_bf.lo.data_ptr() if res_bf.lo is not None else None)
    if out_h16 is not None:  # extra copy of the output as one IEEE fp16 plane (row stride ld_bf)
        A.out_h16 = out_h16.data_ptr()
        if out_bf is None:
            A.ld_bf = ld_bf
    import ctypes
    call('pram_gemm_tc', ctypes.byref(A), stream_ptr())


def conv_tc(x: Split, w: Split, bias: Optional[Tensor], ksize: int, stride: int, relu: bool, split: int,
            res: Optional[Tensor] = None, want_f32: bool = False, want_bf: bool = True, want_ps: bool = False,
            l2norm: bool = False, out_shape_hw=None, bn: int = 0, f16: bool = False, res_bf: Optional[Split] = None,
            want_h16: bool = False):
    """3x3 / 1x1 convolution on tensor cores.  x: Split [B,H,W,Cin] NHWC (stride 1) or the 2x2 phase-split
    tensor [B*4,ceil(H/2),ceil(W/2),Cin] of it (stride 2; then ``out_shape_hw`` = (Ho, Wo) of the conv).
    w: Split [taps,Cout,Cin].  Returns dict with any of 'f32' [B,Ho,Wo,Cout], 'bf' Split, 'ps' Split."""
    dev = x.hi.device
    taps_n, cout, cin = w.shape
    if stride == 1:
        b, h, wd, _ = x.shape
        ho, wo = h, wd
        taps = _S1_TAPS if ksize == 3 else [(0, 0, 0)]
        in_planes, ppi, in_h, in_w = b, 1, h, wd
    else:
        assert ksize == 3 and out_shape_hw is not None
        b4, in_h, in_w, _ = x.shape
        b = b4 // 4
        ho, wo = out_shape_hw
        taps, in_planes, ppi = _S2_TAPS, b4, 4
    out = {}
    f32 = torch.empty((b, ho, wo, cout), device=dev, dtype=torch.float32) if want_f32 else None
    obf = empty_split((b, ho, wo, cout), dev, with_lo=(split == 3)) if want_bf else None
    ops_ps = None
    if want_ps:
        ops_ps = empty_split((b * 4, (ho + 1) // 2, (wo + 1) // 2, cout), dev, with_lo=(split == 3),
                             zero=bool(ho % 2 or wo % 2))
    # 'h16': the output once more as ONE plane of IEEE fp16 bits (bfloat16-typed storage, like as_f16_plane)
    h16 = torch.empty((b, ho, wo, cout), device=dev, dtype=torch.bfloat16) if want_h16 else None
    tw_log2 = 4 if wo >= 16 else max(0, (wo - 1).bit_length())
    gemm_tc(x, x.shape[-1], in_w, in_h, in_planes, cin, w, taps_n, b, ho, wo, cout, taps, ppi, tw_log2, 0, bias, res,
            cout, relu, f32, cout, obf, cout, ops_ps, cout, l2norm, split, bn, f16=f16, res_bf=res_bf, out_h16=h16)
    if h16 is not None:
        out['h16'] = Split(h16, None)
    if f32 is not None:
        out['f32'] = f32
    if obf is not None:
        out['bf'] = obf
    if ops_ps is not None:
        out['ps'] = ops_ps
    return out


def linear_tc(a: Split, lda: int, rows: int, k: int, w: Split, n: int, bias: Optional[Tensor] = None,
              res: Optional[Tensor] = None, ldres: int = 0, relu: bool = False, out_f32: Optional[Tensor] = None,
              ld_f32: int = 0, out_bf: Optional[Split] = None, ld_bf: int = 0, split: int = 3, batch: int = 1,
              w_batched: bool = False, bn: int = 0, qkv: Optional[dict] = None):
    """out[rows, n] = a[rows, k] @ w[n, k]^T on tensor cores (row strides allow column slices of wider
    buffers).  batch > 1: a is [batch, rows, lda], w is [batch, n, k] when ``w_batched``."""
    gemm_tc(a, lda, rows, 1, batch, k, w, (batch if w_batched else 1), batch, 1, rows, n, [(0, 0, 0)], 1, 7,
            1 if w_batched else 0, bias, res, ldres, relu, out_f32, ld_f32, out_bf, ld_bf, None, 0, False, split, bn, qkv)


def mlp_block_tc(a: Split, lda: int, rows: int, w1: Split, w3: Split, tables: Tensor,
                 res: Optional[Tensor], res_ld: int, out_f32: Optional[Tensor], ld_f32: int, out_bf: Optional[Split],
                 ld_bf: int, split: int = 3, dbg: Optional[Tensor] = None):
    """Fused transformer-block tail on tensor cores (csrc/mlp_block_tc.cu): a = [x | ctx] rows [rows, 512] (split bf16,
    row stride ``lda``), w1 [512,512] = mlp.0 with the attention projection folded in, LayerNorm + GELU, w3 [256,512] =
    mlp.3, + residual -> fp32 / split-bf16 rows.  ``tables`` = HOST fp32 [b1 512 | LN gamma 512 | LN beta 512 | b3 256]
    (passed to the kernel as a parameter block).  ``res`` = fp32 residual rows, or None: the residual is x = hi + lo of
    the left half of ``a``.  Reference nets/segnetvit.py:104-106, nets/gml.py:135-137, 182-186."""
    import ctypes
    if tables.is_cuda or tables.dtype != torch.float32 or tables.numel() != 1792 or not tables.is_contiguous():
        raise _lib.PramError('mlp_block_tc: tables must be a contiguous host fp32 tensor of 1792 values')
    A = _lib.MlpBlockArgs()
    A.a_hi, A.a_lo, A.lda, A.T = a.hi.data_ptr(), (a.lo.data_ptr() if a.lo is not None else None), lda, rows
    A.w1_hi, A.w1_lo = w1.hi.data_ptr(), (w1.lo.data_ptr() if w1.lo is not None else None)
    A.w3_hi, A.w3_lo = w3.hi.data_ptr(), (w3.lo.data_ptr() if w3.lo is not None else None)
    A.tables_host = tables.data_ptr()
    A.res, A.res_ld = (res.data_ptr() if res is not None else None), res_ld
    A.out_f32, A.ld_f32 = (out_f32.data_ptr() if out_f32 is not None else None), ld_f32
    if out_bf is not None:
        A.out_hi, A.out_lo, A.ld_bf = out_bf.hi.data_ptr(), (out_bf.lo.data_ptr() if out_bf.lo is not None else None), ld_bf
    A.split = split
    A.dbg = dbg.data_ptr() if dbg is not None else None
    call('pram_mlp_block_tc', ctypes.byref(A), stream_ptr())


CONV1A_TC = True  # tcgen05 kernel (conv1a_tc.cu); False = the CUDA-core kernel (also used when an fp32 copy is wanted)


def conv1a(image_nchw: Tensor, w: Tensor, bias: Tensor, split: int, want_f32: bool = False):
    """conv1a + BN + ReLU from the NCHW fp32 image -> phase-split Split [B*4,ceil(H/2),ceil(W/2),64]
    (and optionally fp32 NHWC)."""
    image_nchw = _f32c(image_nchw)
    b, _, h, wd = image_nchw.shape
    dev = image_nchw.device
    ps = empty_split((b * 4, (h + 1) // 2, (wd + 1) // 2, 64), dev, with_lo=(split == 3), zero=bool(h % 2 or wd % 2))
    f32 = torch.empty((b, h, wd, 64), device=dev, dtype=torch.float32) if want_f32 else None
    if CONV1A_TC and not want_f32:
        call('pram_conv1a_tc', ptr(image_nchw), ptr(w), ptr(bias), b, h, wd, ptr(ps.hi), ptr(ps.lo), split, stream_ptr())
    else:
        call('pram_conv1a', ptr(image_nchw), ptr(w), ptr(bias), b, h, wd, ptr(ps.hi), ptr(ps.lo), ptr(f32), stream_ptr())
    return ps, f32


def gconv3x3_split(x_nhwc: Tensor, w: Tensor, bias: Tensor, relu: bool, split: int) -> Split:
    b, h, wd, c = x_nhwc.shape
    out = empty_split((b, h, wd, c), x_nhwc.device, with_lo=(split == 3))
    call('pram_gconv3x3_split', ptr(x_nhwc), ptr(w), ptr(bias), None, ptr(out.hi), ptr(out.lo), b, h, wd, c // 8,
         int(relu), stream_ptr())
    return out


def gconv3x3_tc(x: Split, w: Tensor, bias: Tensor, relu: bool, split: int) -> Split:
    """Grouped 3x3 conv (groups of 8 channels) on tensor cores: split-bf16 NHWC planes in and out."""
    b, h, wd, c = x.shape
    out = empty_split((b, h, wd, c), x.hi.device, with_lo=(split == 3))
    call('pram_gconv3x3_tc', ptr(x.hi), ptr(x.lo), ptr(w), ptr(bias), ptr(out.hi), ptr(out.lo), b, h, wd, c, int(relu), split,
         stream_ptr())
    return out


def layernorm_gelu_split(x: Tensor, gamma: Tensor, beta: Tensor, c: int, out: Split, gelu: bool = True):
    rows = x.numel() // c
    call('pram_layernorm_gelu_split', ptr(x), ptr(gamma), ptr(beta), None, ptr(out.hi), ptr(out.lo), rows, c, int(gelu),
         stream_ptr())
    return out


# ---- tensor-core flash attention ---------------------------------------------------------------------

ATT_KV_TILE = int(_os.environ.get('PRAM_ATT_KV_TILE', '0'))  # key-tile variant of pram_attention_tc: 0 = auto, 64, 128
# PRAM_ATT_P16=0 forces bf16 hi / lo attention probabilities even where a network asked for the fp16 plane (A/B, bisecting)
ATT_P16_ALLOWED = _os.environ.get('PRAM_ATT_P16', '1') != '0' 


def attention_prep(qkv: Tensor, nparts: int, b: int, n: int, heads: int, cos: Optional[Tensor], sin: Optional[Tensor],
                   scale_qk: float, split: int):
    """qkv rows [b*n, nparts*heads*64] fp32 -> (Q Split [b*heads, n, 64], K Split or None, V^T Split
    [b*heads, 64, n_pad], n_pad)."""
    dev = qkv.device
    lo = split == 3
    n_pad = (n + 7) // 8 * 8
    q = empty_split((b * heads, n, 64), dev, lo)
    k = empty_split((b * heads, n, 64), dev, lo) if nparts == 3 else None
    vt = empty_split((b * heads, 64, n_pad), dev, lo)
    call('pram_attention_prep', ptr(qkv), nparts, b, n, heads, ptr(cos), ptr(sin), float(scale_qk), ptr(q.hi), ptr(q.lo),
         ptr(k.hi) if k is not None else None, ptr(k.lo) if k is not None else None, ptr(vt.hi), ptr(vt.lo), n_pad,
         stream_ptr())
    return q, k, vt, n_pad


def lse_ld(n: int) -> int:
    """Row stride of the attention row-statistics buffers: whole 128-column tiles (read with 16-byte loads)."""
    return (n + 127) // 128 * 128


def attention_tc(q: Split, k: Split, vt: Split, b: int, heads: int, nq: int, nk: int, nk_pad: int, scale: float,
                 out_f32: Optional[Tensor], out_bf: Optional[Split], out_ld: int, split: int, v_mn: bool = False,
                 nk_counts: Optional[Tensor] = None, lse_out: Optional[Tensor] = None, v_f16: bool = False, kv_shift: int = 0):
    """``vt`` is V^T [b*heads, 64, nk_pad] (v_mn=False) or V itself [b*heads, nk, 64] (v_mn=True).  ``nk_counts`` [b] int32:
    keys >= nk_counts[i] of batch element i are padding and receive no attention.  ``lse_out`` [b*heads, lse_ld(nq)] fp32
    (optional) receives the log2-domain log-sum-exp of every query row (input of ``attention_colsum_tc``).  ``v_f16``: the V
    planes hold IEEE fp16 hi / lo (``split_f16`` / the qkv epilogue with ``v_f16``) and P is fed back as one fp16 plane."""
    args = (ptr(q.hi), ptr(q.lo), ptr(k.hi), ptr(k.lo), ptr(vt.hi), ptr(vt.lo), b, heads, nq, nk, nk_pad,
            float(scale), ptr(out_f32), ptr(out_bf.hi) if out_bf is not None else None,
            ptr(out_bf.lo) if (out_bf is not None and out_bf.lo is not None) else None, out_ld, split, ATT_KV_TILE,
            int(v_mn) | (2 if v_f16 else 0), ptr(nk_counts))
    if kv_shift:  # query batch element i attends to keys / values / nk_counts of element (i + kv_shift) mod b
        assert lse_out is None
        call('pram_attention_tc_shift', *args, int(kv_shift), stream_ptr())
    elif lse_out is None:
        call('pram_attention_tc', *args, stream_ptr())
    else:
        call('pram_attention_tc_lse', *args, ptr(lse_out), lse_out.shape[-1], stream_ptr())


def attention_colmean_tc(keys: Split, queries: Split, b: int, heads: int, nkeys: int, nqueries: int, scale: float,
                         lse: Tensor, colsum: Tensor, out: Tensor, out_stride: int, split: int,
                         nq_counts: Optional[Tensor] = None):
    """Mean attention every key receives (over heads and valid queries) -> out[(b*nkeys + j) * out_stride], on tcgen05:
    column sums of softmax(Q K^T) from S^T tiles and the queries' row statistics ``lse`` (``attention_tc(lse_out=...)``),
    then a fixed-order reduction over heads.  ``colsum`` [b*heads, >= nkeys] fp32 scratch.  Reference nets/adagml.py:148, 229."""
    call('pram_attention_colsum_tc', ptr(keys.hi), ptr(keys.lo), ptr(queries.hi), ptr(queries.lo), b, heads, nkeys, nqueries,
         float(scale), ptr(lse), lse.shape[-1], ptr(colsum), colsum.shape[-1], split, ATT_KV_TILE, ptr(nq_counts), stream_ptr())
    call('pram_colmean_reduce', ptr(colsum), colsum.shape[-1], b, heads, nkeys, nqueries, ptr(nq_counts), ptr(out), out_stride,
         stream_ptr())


# ---- K19: batched PnP RANSAC -------------------------------------------------------------------------

def ransac_pnp(kpts: Tensor, matches: Tensor, xyz: Tensor, fx: float, fy: float, cx: float, cy: float,
               max_error: float, pixel_shift: float = 0.5, num_hypotheses: int = 1024, lo_iters: int = 10,
               final_iters: int = 20, min_inliers: int = 3, seed: int = 0):
    """kpts [B,n,2] f32, matches [B,n] i64 (index into xyz or -1), xyz [B,nref,3] f32 ->
    dict(qvec [B,4] wxyz f64, tvec [B,3] f64, num_inliers [B] i32, inliers [B,n] bool, success [B] bool)."""
    _lib.require_cuda(kpts, 'keypoints')
    kpts, xyz = _f32c(kpts), _f32c(xyz)
    matches = matches.contiguous()
    b, n, _ = kpts.shape
    nref = xyz.shape[1]
    dev = kpts.device
    nbytes = int(_lib.load().pram_ransac_workspace_bytes(b, n, num_hypotheses))
    ws = torch.empty((nbytes + 7) // 8, device=dev, dtype=torch.float64)
    q = torch.empty((b, 4), device=dev, dtype=torch.float64)
    t = torch.empty((b, 3), device=dev, dtype=torch.float64)
    ni = torch.empty((b,), device=dev, dtype=torch.int32)
    inl = torch.empty((b, n), device=dev, dtype=torch.uint8)
    ok = torch.empty((b,), device=dev, dtype=torch.int32)
    call('pram_ransac_pnp', ptr(kpts), ptr(matches), ptr(xyz), b, n, nref, float(fx), float(fy), float(cx), float(cy),
         float(pixel_shift), float(max_error), int(num_hypotheses), int(lo_iters), int(final_iters), int(min_inliers),
         int(seed) & 0xffffffff, ptr(ws), ptr(q), ptr(t), ptr(ni), ptr(inl), ptr(ok), stream_ptr())
    return {'qvec': q, 'tvec': t, 'num_inliers': ni, 'inliers': inl.bool(), 'success': ok.bool()}


def ransac_pnp_corr(corr: Tensor, focal_mean: float, max_error: float, num_hypotheses: int = 1024, lo_iters: int = 10,
                    final_iters: int = 20, min_inliers: int = 3, seed: int = 0, counts: Optional[Tensor] = None):
    """corr [B,n,5] float64 (x, y, X, Y, Z), (x, y) = undistorted camera-plane coordinates -> same dict as ``ransac_pnp``.
    Nothing is rounded to float32 on the way in (the pycolmap-compatible entry, localization/pose_estimator.py)."""
    _lib.require_cuda(corr, 'correspondences')
    if corr.dtype != torch.float64:
        raise _lib.PramError('ransac_pnp_corr expects float64 correspondences')
    corr = corr.contiguous()
    b, n, _ = corr.shape
    dev = corr.device
    nbytes = int(_lib.load().pram_ransac_workspace_bytes(b, n, num_hypotheses))
    ws = torch.empty((nbytes + 7) // 8, device=dev, dtype=torch.float64)
    q = torch.empty((b, 4), device=dev, dtype=torch.float64)
    t = torch.empty((b, 3), device=dev, dtype=torch.float64)
    ni = torch.empty((b,), device=dev, dtype=torch.int32)
    inl = torch.empty((b, n), device=dev, dtype=torch.uint8)
    ok = torch.empty((b,), device=dev, dtype=torch.int32)
    call('pram_ransac_pnp_corr', ptr(corr), ptr(counts), b, n, float(focal_mean), float(max_error), int(num_hypotheses),
         int(lo_iters), int(final_iters), int(min_inliers), int(seed) & 0xffffffff, ptr(ws), ptr(q), ptr(t), ptr(ni), ptr(inl),
         ptr(ok), stream_ptr())
    return {'qvec': q, 'tvec': t, 'num_inliers': ni, 'inliers': inl.bool(), 'success': ok.bool()}


# ---- "next" rows: recognition -> matching glue, projection refinement ------------------------------------

def segmentation(logits: Tensor, bg_threshold: float, want_probs: bool = False):
    """logits [T,C] -> (bg_prob [T], seg_id [T] (argmax-1), non_bg [T] bool [, probs [T,C]]);
    reference localization/frame.py:96-121."""
    logits = _f32c(logits)
    t, c = logits.shape
    dev = logits.device
    probs = torch.empty_like(logits) if want_probs else None
    bg = torch.empty((t,), device=dev, dtype=torch.float32)
    sid = torch.empty((t,), device=dev, dtype=torch.int32)
    nb = torch.empty((t,), device=dev, dtype=torch.uint8)
    call('pram_segmentation', ptr(logits), t, c, float(bg_threshold), ptr(probs), ptr(bg), ptr(sid), ptr(nb), stream_ptr())
    return (bg, sid, nb.bool(), probs) if want_probs else (bg, sid, nb.bool())


def rank_landmarks(logits: Tensor, keep: Optional[Tensor], topk: int, max_ranks: int = 8):
    """logits [B,N,C], keep [B,N] bool or None -> dict(sid, rank, count, score [B,topk], n [B],
    label_at_rank [B,max_ranks,N]); reference localization/multimap3d.py:348-379."""
    logits = _f32c(logits)
    b, n, c = logits.shape
    dev = logits.device
    k8 = keep.to(torch.uint8).contiguous() if keep is not None else None
    mk = lambda dt: torch.zeros((b, topk), device=dev, dtype=dt)
    sid, rank, cnt, score = mk(torch.int32), mk(torch.int32), mk(torch.int32), mk(torch.float32)
    ne = torch.empty((b,), device=dev, dtype=torch.int32)
    lab = torch.full((b, max_ranks, n), -1, device=dev, dtype=torch.int32)
    call('pram_rank_landmarks', ptr(logits), ptr(k8), b, n, c, topk, max_ranks, ptr(sid), ptr(rank), ptr(cnt), ptr(score),
         ptr(ne), ptr(lab), stream_ptr())
    return {'sid': sid, 'rank': rank, 'count': cnt, 'score': score, 'n': ne, 'label_at_rank': lab}


def match_by_projection(q_kpts: Tensor, q_descs: Tensor, xyz: Tensor, descs: Tensor, R, t, fx, fy, cx, cy, width, height,
                        threshold: float, ratio: float = 0.995, split: int = 3):
    """K18: q_kpts [M,2], q_descs [M,128], xyz [N,3], descs [N,128] (device fp32), pose (R [3,3], t [3]) ->
    (match [M] i64 index into xyz or -1, d0 [M], d1 [M]).  The M x N similarity runs on the tcgen05 GEMM."""
    dev = q_kpts.device
    m, n = q_kpts.shape[0], xyz.shape[0]
    pose = torch.cat([torch.as_tensor(R, dtype=torch.float64).reshape(9), torch.as_tensor(t, dtype=torch.float64).reshape(3)]).to(dev)
    uv = torch.empty((n, 2), device=dev, dtype=torch.float32)
    valid = torch.empty((n,), device=dev, dtype=torch.uint8)
    call('pram_project_points', ptr(_f32c(xyz)), n, ptr(pose), float(fx), float(fy), float(cx), float(cy), float(width),
         float(height), ptr(uv), ptr(valid), stream_ptr())
    sim = torch.empty((m, n), device=dev, dtype=torch.float32)
    linear_tc(split_bf16(_f32c(q_descs), split == 3), q_descs.shape[1], m, q_descs.shape[1],
              split_bf16(_f32c(descs), split == 3), n, out_f32=sim, ld_f32=n, split=split)
    match = torch.empty((m,), device=dev, dtype=torch.int64)
    d0 = torch.empty((m,), device=dev, dtype=torch.float32)
    d1 = torch.empty((m,), device=dev, dtype=torch.float32)
    call('pram_projection_top2', ptr(sim), n, m, n, ptr(_f32c(q_kpts)), ptr(uv), ptr(valid), float(2 * threshold),
         float(ratio), ptr(match), ptr(d0), ptr(d1), stream_ptr())
    return match, d0, d1


def nearest_neighbor_match(desc0: Tensor, desc1: Tensor, ratio_threshold: Optional[float], distance_threshold: Optional[float],
                           do_mutual_check: bool = True, split: int = 3):
    """desc0 [B,D,N], desc1 [B,D,M] (the reference matcher's 'bdn' layout) -> (matches0 [B,N] i64, scores0 [B,N]);
    reference localization/matchers/nearest_neighbor.py:5-56.  sim = desc0^T desc1 runs on the tcgen05 GEMM."""
    _lib.require_cuda(desc0, 'descriptors0')
    b, d, n = desc0.shape
    m = desc1.shape[2]
    dev = desc0.device
    a0 = split_bf16(_f32c(desc0.transpose(1, 2)), split == 3)  # [B,N,D] K-major rows (layout plumbing only)
    a1 = split_bf16(_f32c(desc1.transpose(1, 2)), split == 3)
    if d % 8:  # 16-byte TMA strides
        raise _lib.PramError('descriptor dimension must be a multiple of 8')
    sim = torch.empty((b, n, m), device=dev, dtype=torch.float32)
    linear_tc(Split(a0.hi.view(b * n, d), a0.lo.view(b * n, d) if a0.lo is not None else None), d, n, d,
              Split(a1.hi.view(b * m, d), a1.lo.view(b * m, d) if a1.lo is not None else None), m, out_f32=sim, ld_f32=m,
              split=split, batch=b, w_batched=True)
    simt = m1 = s1 = None
    if do_mutual_check:
        simt = torch.empty((b, m, n), device=dev, dtype=torch.float32)
        linear_tc(Split(a1.hi.view(b * m, d), a1.lo.view(b * m, d) if a1.lo is not None else None), d, m, d,
                  Split(a0.hi.view(b * n, d), a0.lo.view(b * n, d) if a0.lo is not None else None), n, out_f32=simt, ld_f32=n,
                  split=split, batch=b, w_batched=True)
        m1 = torch.empty((b, m), device=dev, dtype=torch.int64)
        s1 = torch.empty((b, m), device=dev, dtype=torch.float32)
    m0 = torch.empty((b, n), device=dev, dtype=torch.int64)
    s0 = torch.empty((b, n), device=dev, dtype=torch.float32)
    call('pram_nn_match', ptr(sim), ptr(simt), b, n, m, float(ratio_threshold or 0.0), float(distance_threshold or 0.0),
         int(bool(do_mutual_check)), ptr(m0), ptr(s0), ptr(m1), ptr(s1), stream_ptr())
    return m0, s0
