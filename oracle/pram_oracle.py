"""CPU oracle for the PRAM per-frame localization hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pram_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs do, and there only as the checker / the timed CPU baseline.

It is a *functional restatement* (state-dict in, tensors out; torch-CPU fp32 arithmetic, no
``nn.Module``) of the reference algorithms, each function citing the reference ``file:line`` it
follows.  It has to travel to the GPU box, where ``/root/reference`` does not exist, so it imports
nothing from the reference.  It is pinned two ways (tests/test_oracle_pin.py):

* against the reference's own modules imported from ``/root/reference`` (when mounted), tensor for
  tensor on seeded inputs, and
* against golden vectors under ``tests/golden/`` produced by ``oracle/make_golden.py`` from those
  same reference modules (the reference ships no tests / golden vectors of its own).

Parity status: SFD2 / SegNetViT / GML / AdaGML / Sinkhorn / match extraction are PINNED.
``absolute_pose_estimation`` is **parity unpinned**: the arithmetic lives in the third-party wheel
``pycolmap==0.6.1`` (reference ``environment.yml:129``), which is neither under ``/root/reference``
nor installed; the restatement here follows COLMAP 3.9's published algorithm (P3P minimal solver +
LO-RANSAC + non-linear refinement) and is anchored on the reference's call sites only
(``localization/singlemap3d.py:168-175``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-5  # nn.BatchNorm2d default, reference nets/sfd2.py:89

RGB_MEAN = (0.485, 0.456, 0.406)  # reference nets/sfd2.py:14
RGB_STD = (0.229, 0.224, 0.225)  # reference nets/sfd2.py:15


# --------------------------------------------------------------------------------------------
# SFD2 (reference nets/sfd2.py)
# --------------------------------------------------------------------------------------------

def _bn(x: Tensor, sd: Dict[str, Tensor], prefix: str) -> Tensor:
    return F.batch_norm(x, sd[prefix + '.running_mean'], sd[prefix + '.running_var'],
                        sd[prefix + '.weight'], sd[prefix + '.bias'], False, 0.0, BN_EPS)


def _cbr(x: Tensor, sd, name: str, stride: int = 1) -> Tensor:
    """conv3x3(+bias) -> BN(eval) -> ReLU; reference nets/sfd2.py:78-91."""
    y = F.conv2d(x, sd[name + '.0.weight'], sd[name + '.0.bias'], stride=stride, padding=1)
    return F.relu(_bn(y, sd, name + '.1'))


def _resblock(x: Tensor, sd, name: str) -> Tensor:
    """1x1 -> BN -> ReLU -> grouped 3x3 (32 groups) -> BN -> ReLU -> 1x1 -> BN -> +x -> ReLU;
    reference nets/sfd2.py:94-124."""
    y = F.relu(_bn(F.conv2d(x, sd[name + '.conv1.weight']), sd, name + '.bn1'))
    y = F.relu(_bn(F.conv2d(y, sd[name + '.conv2.weight'], padding=1, groups=32), sd, name + '.bn2'))
    y = _bn(F.conv2d(y, sd[name + '.conv3.weight']), sd, name + '.bn3')
    return F.relu(y + x)


def sfd2_trunk(sd: Dict[str, Tensor], image: Tensor) -> Dict[str, Tensor]:
    """The 21-conv stack of ``ResNet4x``; reference nets/sfd2.py:280-293,331-333.

    image: [B,3,H,W] fp32, already ImageNet-normalised by the caller.
    Returns out1b/out2b/out3b/out4, logits [B,65,H/8,W/8] and the L2-normalised desc_map
    [B,128,H/4,W/4] (NCHW).
    """
    out1a = _cbr(image, sd, 'conv1a')
    out1b = _cbr(out1a, sd, 'conv1b', 2)
    out2a = _cbr(out1b, sd, 'conv2a')
    out2b = _cbr(out2a, sd, 'conv2b', 2)
    out3a = _cbr(out2b, sd, 'conv3a')
    out3b = _cbr(out3a, sd, 'conv3b')
    out4 = out3b
    for i in range(3):
        out4 = _resblock(out4, sd, f'conv4.{i}')
    p = F.conv2d(out4, sd['convPa.0.weight'], sd['convPa.0.bias'], stride=2, padding=1)
    p = F.relu(_bn(p, sd, 'convPa.1'))
    p = F.conv2d(p, sd['convPa.3.weight'], sd['convPa.3.bias'], padding=1)
    logits = F.conv2d(p, sd['convPb.weight'], sd['convPb.bias'])
    d = F.conv2d(out4, sd['convDa.0.weight'], sd['convDa.0.bias'], padding=1)
    d = F.relu(_bn(d, sd, 'convDa.1'))
    d = F.conv2d(d, sd['convDa.3.weight'], sd['convDa.3.bias'], padding=1)
    desc = F.conv2d(d, sd['convDb.weight'], sd['convDb.bias'])
    desc = F.normalize(desc, dim=1)
    return {'out1b': out1b, 'out2b': out2b, 'out3b': out3b, 'out4': out4,
            'logits': logits, 'desc_map': desc}


def score_map_from_logits(logits: Tensor, ih: Optional[int] = None, iw: Optional[int] = None) -> Tensor:
    """softmax over 65 channels, drop the dustbin, 8x8 pixel shuffle, optional bilinear resize
    (align_corners=True) when the frame is not a multiple of 8; reference nets/sfd2.py:294-303."""
    b, c, hc, wc = logits.shape
    prob = torch.softmax(logits, dim=1)[:, :-1]  # [B,64,Hc,Wc]
    s = prob.reshape(b, 8, 8, hc, wc).permute(0, 3, 1, 4, 2).reshape(b, hc * 8, wc * 8)
    if ih is not None and (hc * 8 != ih or wc * 8 != iw):
        s = F.interpolate(s[:, None], size=[ih, iw], align_corners=True, mode='bilinear')[:, 0]
    return s


def simple_nms(scores: Tensor, radius: int) -> Tensor:
    """Reference nets/sfd2.py:20-35.  Exact float equality against (2r+1)^2 max-pools whose
    implicit padding behaves as -inf; two suppression/re-detection rounds."""
    k = 2 * radius + 1

    def mp(x):
        return F.max_pool2d(x, kernel_size=k, stride=1, padding=radius)

    zeros = torch.zeros_like(scores)
    keep = scores == mp(scores)
    for _ in range(2):
        supp = mp(keep.float()) > 0
        rest = torch.where(supp, zeros, scores)
        keep = keep | ((rest == mp(rest)) & ~supp)
    return torch.where(keep, scores, zeros)


def select_keypoints(nms: Tensor, conf_th: float, min_keypoints: int, max_keypoints: int,
                     border: int) -> Tuple[Tensor, Tensor]:
    """One frame ([H,W]) of reference nets/sfd2.py:306-329.

    Returns (keypoints [n,2] float32 as (x,y), scores [n]).  Ordering rule: row-major (y,x) when
    n <= max_keypoints, ``torch.topk`` (score-descending) otherwise.  The ``<= min_keypoints``
    fallback to 0.5*conf_th is evaluated on THIS frame (equivalent to a B=1 reference call; the
    reference tests frame 0 of the batch only, nets/sfd2.py:311).
    """
    h, w = nms.shape
    th = np.float32(conf_th)
    yx = torch.nonzero(nms >= float(th))
    if yx.shape[0] <= min_keypoints:
        yx = torch.nonzero(nms >= float(conf_th * 0.5))
    sc = nms[yx[:, 0], yx[:, 1]]
    m = (yx[:, 0] >= border) & (yx[:, 0] < h - border) & (yx[:, 1] >= border) & (yx[:, 1] < w - border)
    yx, sc = yx[m], sc[m]
    if 0 <= max_keypoints < yx.shape[0]:
        sc, idx = torch.topk(sc, max_keypoints, dim=0)
        yx = yx[idx]
    return torch.flip(yx, [1]).float(), sc


def sample_map(kpts: Tensor, fmap: Tensor, s: int = 4, normalize: bool = True) -> Tensor:
    """Bilinear sampling of a [1,C,h,w] map at pixel keypoints [n,2] (x,y);
    reference nets/sfd2.py:53-64 (descriptors) and :348-363 (mid features).  Returns [C,n]."""
    _, c, h, w = fmap.shape
    g = kpts - s / 2 + 0.5
    g = g / torch.tensor([w * s - s / 2 - 0.5, h * s - s / 2 - 0.5]).to(g)[None]
    g = g * 2 - 1
    out = F.grid_sample(fmap, g.view(1, 1, -1, 2), mode='bilinear', align_corners=True).reshape(1, c, -1)
    if normalize:
        out = F.normalize(out, p=2, dim=1)
    return out[0]


def sfd2_extract_local_global(sd, image: Tensor, config: Optional[dict] = None) -> dict:
    """Reference nets/sfd2.py:269-346 for a [B,3,H,W] batch, frame-by-frame selection."""
    cfg = {'conf_th': 0.005, 'remove_borders': 4, 'min_keypoints': 128, 'max_keypoints': 4096}
    cfg.update(config or {})
    b, _, ih, iw = image.shape
    t = sfd2_trunk(sd, image)
    score = score_map_from_logits(t['logits'], ih, iw)
    nms = simple_nms(score, 4)
    kps, scs, descs = [], [], []
    for i in range(b):
        k, s = select_keypoints(nms[i], cfg['conf_th'], cfg['min_keypoints'], cfg['max_keypoints'],
                                cfg['remove_borders'])
        kps.append(k)
        scs.append(s)
        descs.append(sample_map(k, t['desc_map'][i:i + 1], 4, True))
    return {'score_map': score, 'desc_map': t['desc_map'], 'mid_features': t['out4'],
            'global_descriptors': [t['out1b'], t['out2b'], t['out3b'], t['out4']],
            'keypoints': kps, 'scores': tuple(scs), 'descriptors': descs, 'logits': t['logits'],
            'nms': nms}


def sfd2_sample(score_map: Tensor, fmap: Tensor, kpts: Tensor, s: int = 4, norm_desc: bool = True):
    """Reference nets/sfd2.py:348-369: (scores [n], descriptors [C,n])."""
    d = sample_map(kpts, fmap, s, norm_desc)
    sc = score_map[0, kpts[:, 1].long(), kpts[:, 0].long()]
    return sc, d


def sfd2_extract_return(sd, img: Tensor, conf_th: float = 0.001, topK: int = -1,
                        scales: Sequence[float] = (1.0,)) -> Optional[dict]:
    """The offline-export variant, reference nets/sfd2.py:386-589 (mask=None branch), without its
    hard-coded ``.cuda()`` calls (:396,:476).  img: [1,3,H,W] in [0,1], NOT normalised.
    NMS radius 3, strict ``>`` threshold, score-descending stable order, border 4, the *simple*
    sampling normalisation x/(w/2)-1, float64 outputs."""
    mean = torch.tensor(RGB_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(RGB_STD).view(1, 3, 1, 1)
    x = (img.reshape(1, 3, img.shape[-2], img.shape[-1]) - mean) / std
    _, _, H, W = x.shape
    pts_all, desc_all = [], []
    for s in scales:
        xi = x if s == 1.0 else F.interpolate(x, size=(int(H * s), int(W * s)), mode='bilinear',
                                              align_corners=True)
        nh, nw = xi.shape[2:]
        t = sfd2_trunk(sd, xi)
        heat = score_map_from_logits(t['logits'], nh, nw)
        nms = simple_nms(heat[:, None], 3)[0, 0]
        yx = torch.nonzero(nms > conf_th)
        sc = nms[yx[:, 0], yx[:, 1]].numpy()
        xy = torch.flip(yx, [1]).float().numpy()
        order = np.argsort(sc)[::-1]
        xy, sc = xy[order], sc[order]
        ok = ~((xy[:, 0] < 4) | (xy[:, 0] >= W - 4) | (xy[:, 1] < 4) | (xy[:, 1] >= H - 4))
        xy, sc = xy[ok], sc[ok]
        if xy.shape[0] == 0:
            continue
        g = torch.from_numpy(xy.copy())
        g[:, 0] = g[:, 0] / (float(nw) / 2.) - 1.
        g[:, 1] = g[:, 1] / (float(nh) / 2.) - 1.
        d = F.grid_sample(t['desc_map'], g.view(1, 1, -1, 2).float(), mode='bilinear',
                          align_corners=True).numpy().reshape(t['desc_map'].shape[1], -1)
        d = d / np.linalg.norm(d, axis=0)[None]
        xy[:, 0] = xy[:, 0] * W / nw
        xy[:, 1] = xy[:, 1] * H / nh
        pts_all.append(np.concatenate([xy, sc[:, None]], 1))
        desc_all.append(d.T)
    if not pts_all:
        return None
    pts = np.vstack(pts_all)
    desc = np.vstack(desc_all)
    kp, sc = pts[:, :2], pts[:, 2]
    if topK > 0:
        idx = np.array(sc, dtype=float).argsort()[::-1][:topK]
        kp, sc, desc = kp[idx], sc[idx], desc[idx]
    return {'keypoints': np.array(kp, dtype=float), 'descriptors': np.array(desc, dtype=float),
            'scores': np.array(sc, dtype=float)}


# --------------------------------------------------------------------------------------------
# shared transformer pieces (reference nets/segnetvit.py, nets/gml.py, nets/utils.py)
# --------------------------------------------------------------------------------------------

def normalize_keypoints(kpts: Tensor, image_shape) -> Tensor:
    """(k - (W,H)/2) / (0.7*max(W,H)) with ``_,_,height,width = image_shape``;
    reference nets/utils.py:17-24."""
    _, _, height, width = image_shape
    size = torch.tensor([float(width), float(height)]).to(kpts)
    return (kpts - size / 2) / (size.max() * 0.7)


def fourier_encoding(wr: Tensor, nkpts: Tensor) -> Tensor:
    """Learnable Fourier positional encoding -> [2,B,1,N,64] (cos | sin, each value repeated twice
    along the last axis); reference nets/segnetvit.py:35-40 == nets/gml.py:69-74."""
    proj = nkpts @ wr.t()  # [B,N,32]
    emb = torch.stack([torch.cos(proj), torch.sin(proj)], 0).unsqueeze(-3)
    return emb.repeat_interleave(2, dim=-1)


def _rotary(enc: Tensor, t: Tensor) -> Tensor:
    """t*cos + rotate_half(t)*sin on ADJACENT pairs (2i,2i+1); reference nets/segnetvit.py:15-23."""
    tp = t.unflatten(-1, (-1, 2))
    rot = torch.stack((-tp[..., 1], tp[..., 0]), dim=-1).flatten(-2)
    return t * enc[0] + rot * enc[1]


def _mlp(sd, pre: str, x: Tensor) -> Tensor:
    """Linear -> LayerNorm -> GELU(erf) -> Linear; reference nets/segnetvit.py:90-95."""
    y = F.linear(x, sd[pre + '.0.weight'], sd[pre + '.0.bias'])
    y = F.layer_norm(y, (y.shape[-1],), sd[pre + '.1.weight'], sd[pre + '.1.bias'])
    y = F.gelu(y)
    return F.linear(y, sd[pre + '.3.weight'], sd[pre + '.3.bias'])


def self_block(sd, pre: str, x: Tensor, enc: Optional[Tensor], heads: int = 4,
               return_attn_mean: bool = False):
    """One SelfMultiHeadAttention block; reference nets/segnetvit.py:97-106 == nets/gml.py:128-137.
    qkv output features are interleaved (head, dim, {q,k,v}) innermost."""
    b, n, _ = x.shape
    qkv = F.linear(x, sd[pre + '.qkv.weight'], sd[pre + '.qkv.bias'])
    qkv = qkv.reshape(b, n, heads, -1, 3).transpose(1, 2)  # [B,h,N,64,3]
    q, k, v = qkv[..., 0], qkv[..., 1], qkv[..., 2]
    if enc is not None:
        q, k = _rotary(enc, q), _rotary(enc, k)
    attn = torch.softmax(torch.einsum('bhid,bhjd->bhij', q, k) * (q.shape[-1] ** -0.5), -1)
    ctx = torch.einsum('bhij,bhjd->bhid', attn, v)
    msg = F.linear(ctx.transpose(1, 2).flatten(-2), sd[pre + '.proj.weight'], sd[pre + '.proj.bias'])
    out = x + _mlp(sd, pre + '.mlp', torch.cat([x, msg], -1))
    if return_attn_mean:  # reference nets/adagml.py:148
        return out, attn.mean(1).mean(1)
    return out


def cross_block(sd, pre: str, x0: Tensor, x1: Tensor, heads: int = 4, return_attn_mean: bool = False):
    """Bidirectional cross attention with shared qk projection; reference nets/gml.py:164-186."""
    def split(t):
        return t.unflatten(-1, (heads, -1)).transpose(1, 2)

    qk0 = split(F.linear(x0, sd[pre + '.to_qk.weight'], sd[pre + '.to_qk.bias']))
    qk1 = split(F.linear(x1, sd[pre + '.to_qk.weight'], sd[pre + '.to_qk.bias']))
    v0 = split(F.linear(x0, sd[pre + '.to_v.weight'], sd[pre + '.to_v.bias']))
    v1 = split(F.linear(x1, sd[pre + '.to_v.weight'], sd[pre + '.to_v.bias']))
    sc = (qk0.shape[-1] ** -0.5) ** 0.5
    sim = torch.einsum('bhid,bhjd->bhij', qk0 * sc, qk1 * sc)
    a01 = torch.softmax(sim, -1)
    a10 = torch.softmax(sim.transpose(-2, -1), -1)
    m0 = torch.einsum('bhij,bhjd->bhid', a01, v1).transpose(1, 2).flatten(-2)
    m1 = torch.einsum('bhij,bhjd->bhid', a10, v0).transpose(1, 2).flatten(-2)
    m0 = F.linear(m0, sd[pre + '.proj.weight'], sd[pre + '.proj.bias'])
    m1 = F.linear(m1, sd[pre + '.proj.weight'], sd[pre + '.proj.bias'])
    y0 = x0 + _mlp(sd, pre + '.mlp', torch.cat([x0, m0], -1))
    y1 = x1 + _mlp(sd, pre + '.mlp', torch.cat([x1, m1], -1))
    if return_attn_mean:  # reference nets/adagml.py:229 (note the order: attn10 first)
        return y0, y1, a10.mean(1).mean(1), a01.mean(1).mean(1)
    return y0, y1


def segnetvit_forward(sd, seg_descriptors: Tensor, keypoints: Tensor, image_shape=None,
                      norm_keypoints: Optional[Tensor] = None, n_layers: int = 15) -> Tensor:
    """Reference nets/segnetvit.py:174-203 -> raw logits [B,N,n_class]."""
    if norm_keypoints is None:
        if image_shape is None:
            raise ValueError('Require image shape for keypoint coordinate normalization')
        norm_keypoints = normalize_keypoints(keypoints, image_shape)
    enc = fourier_encoding(sd['kenc.Wr.weight'], norm_keypoints)
    x = F.linear(seg_descriptors, sd['input_proj.weight'], sd['input_proj.bias'])
    for i in range(n_layers):
        x = self_block(sd, f'gnn.layers.{i}', x, enc)
    return _mlp(sd, 'seg', x)


# --------------------------------------------------------------------------------------------
# GML / AdaGML (reference nets/gml.py, nets/adagml.py)
# --------------------------------------------------------------------------------------------

def sinkhorn_with_dustbin(dist: Tensor, bin_score: Tensor, iters: int) -> Tensor:
    """Reference nets/gml.py:27-46: dustbin row+col, row-softmax, ``iters`` x {u = r/((p*v).sum(-1)
    +1e-8); v = c/((p*u).sum(-2)+1e-8)}, p*u*v.  r = [1..1, M+1], c = [1..1, N+1]."""
    b, m, n = dist.shape
    Z = torch.cat([dist, bin_score.expand(b, m, 1)], -1)
    Z = torch.cat([Z, bin_score.expand(b, 1, n + 1)], -2)
    r = torch.ones(b, m + 1)
    r[:, -1] = m + 1
    c = torch.ones(b, n + 1)
    c[:, -1] = n + 1
    p = torch.softmax(Z, -1)
    u, v = torch.ones_like(r), torch.ones_like(c)
    for _ in range(iters):
        u = r / ((p * v.unsqueeze(-2)).sum(-1) + 1e-8)
        v = c / ((p * u.unsqueeze(-1)).sum(-2) + 1e-8)
    return p * u.unsqueeze(-1) * v.unsqueeze(-2)


def compute_matches(P: Tensor, th: float = 0.2):
    """Mutual arg-max over P[:, :-1, :-1] with threshold; reference nets/gml.py:304-319."""
    inner = P[:, :-1, :-1]
    mx0, mx1 = inner.max(2), inner.max(1)
    i0, i1 = mx0.indices, mx1.indices
    ar0 = torch.arange(i0.shape[1])[None]
    ar1 = torch.arange(i1.shape[1])[None]
    mut0 = ar0 == i1.gather(1, i0)
    mut1 = ar1 == i0.gather(1, i1)
    zero = P.new_tensor(0)
    s0 = torch.where(mut0, mx0.values, zero)
    s1 = torch.where(mut1, s0.gather(1, i1), zero)
    val0 = mut0 & (s0 > th)
    val1 = mut1 & val0.gather(1, i1)
    return (torch.where(val0, i0, i0.new_tensor(-1)), torch.where(val1, i1, i1.new_tensor(-1)), s0, s1)


def _gml_norm_kpts(data: dict):
    if 'norm_keypoints0' in data and 'norm_keypoints1' in data:
        return data['norm_keypoints0'], data['norm_keypoints1']
    if 'image0' in data and 'image1' in data:
        return (normalize_keypoints(data['keypoints0'], data['image0'].shape).float(),
                normalize_keypoints(data['keypoints1'], data['image1'].shape).float())
    if 'image_shape0' in data and 'image_shape1' in data:
        return (normalize_keypoints(data['keypoints0'], data['image_shape0']).float(),
                normalize_keypoints(data['keypoints1'], data['image_shape1']).float())
    raise ValueError('Require image shape for keypoint coordinate normalization')


def gml_forward(sd, data: dict, n_layers: int = 9, sinkhorn_iterations: int = 20, p: float = 0.2,
                return_intermediate: bool = False) -> dict:
    """Reference nets/gml.py:250-294 (``GML.produce_matches``)."""
    nk0, nk1 = _gml_norm_kpts(data)
    d0 = F.linear(data['descriptors0'], sd['input_proj.weight'], sd['input_proj.bias'])
    d1 = F.linear(data['descriptors1'], sd['input_proj.weight'], sd['input_proj.bias'])
    e0 = fourier_encoding(sd['poseenc.Wr.weight'], nk0)
    e1 = fourier_encoding(sd['poseenc.Wr.weight'], nk1)
    for i in range(n_layers):
        d0 = self_block(sd, f'self_attn.{i}', d0, e0)
        d1 = self_block(sd, f'self_attn.{i}', d1, e1)
        d0, d1 = cross_block(sd, f'cross_attn.{i}', d0, d1)
    last = n_layers - 1
    dim = d0.shape[-1]
    m0 = F.linear(d0, sd[f'out_proj.{last}.weight'], sd[f'out_proj.{last}.bias']) / dim ** .25
    m1 = F.linear(d1, sd[f'out_proj.{last}.weight'], sd[f'out_proj.{last}.bias']) / dim ** .25
    dist = torch.einsum('bmd,bnd->bmn', m0, m1)
    P = sinkhorn_with_dustbin(dist, sd['bin_score'], sinkhorn_iterations)
    i0, i1, s0, s1 = compute_matches(P, p)
    out = {'matches0': i0, 'matches1': i1, 'matching_scores0': s0, 'matching_scores1': s1}
    if return_intermediate:
        out.update({'dist': dist, 'P': P, 'desc0': d0, 'desc1': d1})
    return out


def _pooling(sd, pre: str, x: Tensor, score: Tensor) -> Tensor:
    """``PoolingLayer``; reference nets/adagml.py:114-138 -> confidence [B,N]."""
    s = _mlp(sd, pre + '.score_enc', score)
    xp = F.linear(x, sd[pre + '.proj.weight'], sd[pre + '.proj.bias'])
    return torch.sigmoid(_mlp(sd, pre + '.predict', torch.cat([xp, s], -1))).squeeze(-1)


def adagml_conf_threshold(layer: int, n_layers: int = 9) -> float:
    """0.5 + 0.1*exp(-4*l/L); reference nets/adagml.py:516-520."""
    return float(np.clip(0.5 + 0.1 * np.exp(-4.0 * layer / n_layers), 0, 1))


def adagml_forward(sd, data: dict, n_layers: int = 9, n_min_tokens: int = 256,
                   sinkhorn_iterations: int = 20, p: float = 0.2, return_trace: bool = False) -> dict:
    """Reference nets/adagml.py:307-404 (``AdaGML.produce_matches``), B must be 1.
    Uses the device-agnostic Sinkhorn (same maths as nets/adagml.py:42-50, which hard-codes 'cuda')."""
    nk0, nk1 = _gml_norm_kpts(data)
    d0 = F.linear(data['descriptors0'], sd['input_proj.weight'], sd['input_proj.bias'])
    d1 = F.linear(data['descriptors1'], sd['input_proj.weight'], sd['input_proj.bias'])
    e0 = fourier_encoding(sd['poseenc.Wr.weight'], nk0)
    e1 = fourier_encoding(sd['poseenc.Wr.weight'], nk1)
    nb, m, n = d0.shape[0], d0.shape[1], d1.shape[1]
    ind0 = torch.arange(m)[None]
    ind1 = torch.arange(n)[None]
    trace = []
    ni = 0
    for ni in range(n_layers):
        d0, a00 = self_block(sd, f'self_attn.{ni}', d0, e0, return_attn_mean=True)
        d1, a11 = self_block(sd, f'self_attn.{ni}', d1, e1, return_attn_mean=True)
        d0, d1, a01, a10 = cross_block(sd, f'cross_attn.{ni}', d0, d1, return_attn_mean=True)
        c0 = _pooling(sd, f'pooling.{ni}', d0, torch.stack([a00, a01], -1))
        c1 = _pooling(sd, f'pooling.{ni}', d1, torch.stack([a11, a10], -1))
        if ni >= 1:
            th = adagml_conf_threshold(ni, n_layers)
            if d0.shape[1] >= n_min_tokens:
                k0 = c0 > th
                ind0, d0, e0 = ind0[k0][None], d0[k0][None], e0[:, :, k0][:, None]
            if d1.shape[1] >= n_min_tokens:
                k1 = c1 > th
                ind1, d1, e1 = ind1[k1][None], d1[k1][None], e1[:, :, k1][:, None]
            trace.append((ni, d0.shape[1], d1.shape[1]))
            conf = torch.cat([c0, c1], -1)
            pos = 1.0 - (conf < th).float().sum() / (m + n)
            if pos > 0.95:
                break
    dim = d0.shape[-1]
    m0 = F.linear(d0, sd[f'out_proj.{ni}.weight'], sd[f'out_proj.{ni}.bias']) / dim ** .25
    m1 = F.linear(d1, sd[f'out_proj.{ni}.weight'], sd[f'out_proj.{ni}.bias']) / dim ** .25
    dist = torch.einsum('bmd,bnd->bmn', m0, m1)
    P = sinkhorn_with_dustbin(dist, sd['bin_score'], sinkhorn_iterations)
    i0, _, s0, _ = compute_matches(P, p)
    valid = i0 > -1
    mi0 = torch.where(valid)[1]
    mi1 = i0[valid]
    full_i = torch.full((nb, m), -1, dtype=i0.dtype)
    full_i[:, ind0[0, mi0]] = ind1[0, mi1]
    full_s = torch.zeros((nb, m))
    full_s[:, ind0] = s0
    out = {'matches0': full_i, 'matching_scores0': full_s}
    if return_trace:
        out['trace'] = trace
        out['last_layer'] = ni
    return out


# --------------------------------------------------------------------------------------------
# Absolute pose: P3P + LO-RANSAC + refinement.   *** parity unpinned *** (see module docstring)
# Replaces the reference's calls to pycolmap.absolute_pose_estimation
# (localization/singlemap3d.py:168-175, :324, :454; tracker.py:211; pose_estimator.py:213,338,452).
# --------------------------------------------------------------------------------------------

from benchdata import quat_to_rotmat, rotmat_to_quat, pose_error  # noqa: E402,F401  (shared metric helpers)


def p3p_solve(x: np.ndarray, X: np.ndarray) -> List[Tuple[np.ndarray, np.ndarray]]:
    """Minimal absolute pose from 3 bearing/point pairs.  x: [3,2] normalised image points,
    X: [3,3] world points.  Classic Grunert/Fischler-Bolles quartic in the distance ratio, followed
    by a 3-point rigid alignment (the formulation COLMAP 3.9's P3PEstimator uses, Gao et al. 2003).
    Returns a list of (R, t) with X_cam = R X + t."""
    f = np.concatenate([x, np.ones((3, 1))], 1)
    f = f / np.linalg.norm(f, axis=1, keepdims=True)
    a = np.linalg.norm(X[1] - X[2])
    b = np.linalg.norm(X[0] - X[2])
    c = np.linalg.norm(X[0] - X[1])
    if min(a, b, c) < 1e-12:
        return []
    ca, cb, cg = f[1] @ f[2], f[0] @ f[2], f[0] @ f[1]
    a2, b2, c2 = a * a, b * b, c * c
    q = (a2 - c2) / b2
    p = (a2 + c2) / b2
    A4 = (q - 1) ** 2 - 4 * c2 / b2 * ca * ca
    A3 = 4 * (q * (1 - q) * cb - (1 - p) * ca * cg + 2 * c2 / b2 * ca * ca * cb)
    A2 = 2 * (q * q - 1 + 2 * q * q * cb * cb + 2 * (b2 - c2) / b2 * ca * ca
              - 4 * p * ca * cb * cg + 2 * (b2 - a2) / b2 * cg * cg)
    A1 = 4 * (-q * (1 + q) * cb + 2 * a2 / b2 * cg * cg * cb - (1 - p) * ca * cg)
    A0 = (1 + q) ** 2 - 4 * a2 / b2 * cg * cg
    roots = np.roots([A4, A3, A2, A1, A0]) if abs(A4) > 1e-14 else np.roots([A3, A2, A1, A0])
    sols = []
    for r in roots:
        if abs(r.imag) > 1e-8 * max(1.0, abs(r.real)):
            continue
        v = r.real
        if v <= 0:
            continue
        den = 2 * (cg - v * ca)
        if abs(den) < 1e-12:
            continue
        u = ((-1 + q) * v * v - 2 * q * cb * v + 1 + q) / den
        if u <= 0:
            continue
        s1sq = c2 / (1 + u * u - 2 * u * cg)
        if s1sq <= 0:
            continue
        s1 = math.sqrt(s1sq)
        s2, s3 = u * s1, v * s1
        Y = np.stack([s1 * f[0], s2 * f[1], s3 * f[2]])
        # rigid alignment X -> Y from 3 points
        mx, my = X.mean(0), Y.mean(0)
        Hm = (X - mx).T @ (Y - my)
        # add the normal direction so that the 3-point problem is well conditioned
        nx = np.cross(X[1] - X[0], X[2] - X[0])
        ny = np.cross(Y[1] - Y[0], Y[2] - Y[0])
        Hm = Hm + np.outer(nx, ny) / max(np.linalg.norm(nx), 1e-30)
        U, _, Vt = np.linalg.svd(Hm)
        D = np.diag([1, 1, np.sign(np.linalg.det(Vt.T @ U.T))])
        R = Vt.T @ D @ U.T
        t = my - R @ mx
        sols.append((R, t))
    return sols


def _reproj_sq_err(R, t, x, X) -> np.ndarray:
    Xc = X @ R.T + t
    z = Xc[:, 2]
    ok = z > 1e-12
    pr = Xc[:, :2] / np.where(ok, z, 1.0)[:, None]
    e = ((pr - x) ** 2).sum(1)
    return np.where(ok, e, np.inf)


def refine_pose(R, t, x, X, iters: int = 100, loss_scale: Optional[float] = None):
    """Gauss-Newton / LM on the reprojection error in normalised coordinates over (so3, t) with an
    optional Cauchy loss (COLMAP RefineAbsolutePose: Cauchy, scale 1 px [recalled, unverified])."""
    R, t = R.copy(), t.copy()
    lam = 1e-4
    def cost_and_sys(R, t):
        Xc = X @ R.T + t
        z = np.maximum(Xc[:, 2], 1e-9)
        r = np.stack([Xc[:, 0] / z - x[:, 0], Xc[:, 1] / z - x[:, 1]], 1)  # [n,2]
        e2 = (r ** 2).sum(1)
        if loss_scale is not None:
            w = 1.0 / (1.0 + e2 / loss_scale ** 2)
            cost = (loss_scale ** 2 * np.log1p(e2 / loss_scale ** 2)).sum()
        else:
            w = np.ones_like(e2)
            cost = e2.sum()
        n = X.shape[0]
        J = np.zeros((n, 2, 6))
        iz = 1.0 / z
        # d(proj)/d(Xc)
        dx = np.stack([iz, np.zeros(n), -Xc[:, 0] * iz * iz], 1)
        dy = np.stack([np.zeros(n), iz, -Xc[:, 1] * iz * iz], 1)
        # d(Xc)/d(omega) = -[Xc_rot]_x where Xc_rot = R X ; d(Xc)/dt = I
        Xr = Xc - t
        def skew_mul(d):  # d^T * (-[Xr]_x)  -> [n,3]
            return np.cross(Xr, d)
        J[:, 0, :3] = skew_mul(dx)
        J[:, 1, :3] = skew_mul(dy)
        J[:, 0, 3:] = dx
        J[:, 1, 3:] = dy
        Jw = J * w[:, None, None]
        H = np.einsum('nij,nik->jk', Jw, J)
        g = np.einsum('nij,ni->j', Jw, r)
        return cost, H, g
    cost, H, g = cost_and_sys(R, t)
    for _ in range(iters):
        try:
            d = np.linalg.solve(H + lam * np.diag(np.diag(H) + 1e-12), -g)
        except np.linalg.LinAlgError:
            break
        th = np.linalg.norm(d[:3])
        if th > 1e-15:
            k = d[:3] / th
            Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
            dR = np.eye(3) + math.sin(th) * Kx + (1 - math.cos(th)) * Kx @ Kx
        else:
            dR = np.eye(3)
        Rn, tn = dR @ R, t + d[3:]  # X_cam' = exp(w) R X + t + dt
        cn, Hn, gn = cost_and_sys(Rn, tn)
        if cn < cost:
            small = (cost - cn) < 1e-14 * max(cost, 1e-30)
            R, t, cost, H, g = Rn, tn, cn, Hn, gn
            lam = max(lam * 0.3, 1e-12)
            if small or np.linalg.norm(d) < 1e-12:
                break
        else:
            lam *= 10
            if lam > 1e12:
                break
    return R, t


def absolute_pose_estimation(points2D: np.ndarray, points3D: np.ndarray, camera: dict,
                             max_error: float = 12.0, min_num_trials: int = 1000,
                             max_num_trials: int = 100000, confidence: float = 0.9999,
                             min_inlier_ratio: float = 0.01, seed: int = 0) -> Optional[dict]:
    """CPU restatement of the call the reference makes to pycolmap (see section header).

    camera: {'model': 'PINHOLE'|'SIMPLE_PINHOLE'|..., 'width','height','params'} with
    params = (fx,fy,cx,cy[,k1,k2,p1,p2]) or (f,cx,cy[,k1[,k2]]); pixels go to the camera plane through the lens model
    (``cam_from_img``) as COLMAP does.  Parity with pycolmap itself is unpinned.
    Residual = squared reprojection error in normalised camera coordinates, threshold
    (max_error / mean focal)^2 (COLMAP convention).  Returns None on failure, else a dict with
    'qvec' (wxyz), 'tvec', 'num_inliers', 'inliers' (bool[n]).
    """
    p2 = np.asarray(points2D, np.float64)
    p3 = np.asarray(points3D, np.float64)
    n = p2.shape[0]
    if n < 3:
        return None
    fx, fy, cx, cy = camera_intrinsics(camera)
    x = cam_from_img(camera, p2)
    thr = (max_error / (0.5 * (fx + fy))) ** 2
    rng = np.random.RandomState(seed)
    best = (-1, np.inf, None, None)  # inliers, residual sum, R, t
    max_trials = max_num_trials
    trial = 0
    while trial < max_trials:
        trial += 1
        idx = rng.choice(n, 3, replace=False)
        for R, t in p3p_solve(x[idx], p3[idx]):
            e = _reproj_sq_err(R, t, x, p3)
            inl = e <= thr
            cnt = int(inl.sum())
            rs = float(e[inl].sum())
            if cnt > best[0] or (cnt == best[0] and rs < best[1]):
                # local optimisation on the current inlier set
                if cnt >= 6:
                    R2, t2 = refine_pose(R, t, x[inl], p3[inl], iters=10)
                    e2 = _reproj_sq_err(R2, t2, x, p3)
                    inl2 = e2 <= thr
                    if int(inl2.sum()) >= cnt:
                        R, t, cnt, rs = R2, t2, int(inl2.sum()), float(e2[inl2].sum())
                best = (cnt, rs, R, t)
                ratio = cnt / n
                if ratio > 0:
                    pn = 1 - ratio ** 3
                    need = math.inf if pn >= 1 else (0 if pn <= 0 else math.log(1 - confidence) / math.log(pn))
                    max_trials = min(max_num_trials, max(min_num_trials, int(math.ceil(need))))
    if best[0] < 3 or best[0] < min_inlier_ratio * n or best[2] is None:
        return None
    R, t = best[2], best[3]
    inl = _reproj_sq_err(R, t, x, p3) <= thr
    if inl.sum() >= 3:
        R, t = refine_pose(R, t, x[inl], p3[inl], iters=100, loss_scale=1.0 / (0.5 * (fx + fy)))
    inl = _reproj_sq_err(R, t, x, p3) <= thr
    return {'qvec': rotmat_to_quat(R), 'tvec': t, 'num_inliers': int(inl.sum()), 'inliers': inl,
            'R': R}


def camera_intrinsics(camera) -> Tuple[float, float, float, float]:
    """(fx,fy,cx,cy) from a COLMAP-style camera dict / namedtuple (reference
    localization/camera.py:1-11: Camera(id, model, width, height, params))."""
    model = camera['model'] if isinstance(camera, dict) else camera.model
    model = getattr(model, 'name', model)
    params = camera['params'] if isinstance(camera, dict) else camera.params
    params = [float(v) for v in params]
    if model in ('PINHOLE', 'OPENCV', 'FULL_OPENCV', 'OPENCV_FISHEYE'):
        return params[0], params[1], params[2], params[3]
    return params[0], params[0], params[1], params[2]  # SIMPLE_PINHOLE / SIMPLE_RADIAL / RADIAL


def _distort(model: str, extra, u: np.ndarray, v: np.ndarray):
    """COLMAP sensor/models.h ``Distortion`` of SIMPLE_RADIAL / RADIAL / OPENCV: (du, dv) at camera-plane (u, v)."""
    r2 = u * u + v * v
    if model == 'SIMPLE_RADIAL':
        rad = extra[0] * r2
        return u * rad, v * rad
    if model == 'RADIAL':
        rad = extra[0] * r2 + extra[1] * r2 * r2
        return u * rad, v * rad
    if model == 'OPENCV':
        k1, k2, p1, p2 = extra[:4]
        rad = k1 * r2 + k2 * r2 * r2
        return (u * rad + 2 * p1 * u * v + p2 * (r2 + 2 * u * u), v * rad + 2 * p2 * u * v + p1 * (r2 + 2 * v * v))
    return np.zeros_like(u), np.zeros_like(v)


def img_from_cam(camera, uv: np.ndarray) -> np.ndarray:
    """Camera plane -> pixels with the lens model (COLMAP ``ImgFromCam``); used to synthesise distorted observations."""
    model = camera['model'] if isinstance(camera, dict) else camera.model
    params = [float(v) for v in (camera['params'] if isinstance(camera, dict) else camera.params)]
    fx, fy, cx, cy = camera_intrinsics(camera)
    extra = params[3:] if model in ('SIMPLE_RADIAL', 'RADIAL') else params[4:]
    du, dv = _distort(model, extra, uv[:, 0], uv[:, 1])
    return np.stack([fx * (uv[:, 0] + du) + cx, fy * (uv[:, 1] + dv) + cy], 1)


def cam_from_img(camera, xy: np.ndarray) -> np.ndarray:
    """Pixels -> undistorted camera plane: COLMAP's ``IterativeUndistortion`` -- Newton's method on
    f(x) = x + d(x) - x0 with a central-difference Jacobian (step 1e-10 relative, at most 100 iterations)."""
    model = camera['model'] if isinstance(camera, dict) else camera.model
    params = [float(v) for v in (camera['params'] if isinstance(camera, dict) else camera.params)]
    fx, fy, cx, cy = camera_intrinsics(camera)
    p = np.asarray(xy, np.float64)
    u0, v0 = (p[:, 0] - cx) / fx, (p[:, 1] - cy) / fy
    if model in ('SIMPLE_PINHOLE', 'PINHOLE'):
        return np.stack([u0, v0], 1)
    extra = params[3:] if model in ('SIMPLE_RADIAL', 'RADIAL') else params[4:]
    u, v = u0.copy(), v0.copy()
    for _ in range(100):
        eu, ev = np.maximum(1e-10, np.abs(1e-10 * u)), np.maximum(1e-10, np.abs(1e-10 * v))
        du, dv = _distort(model, extra, u, v)
        a = _distort(model, extra, u - eu, v); b = _distort(model, extra, u + eu, v)
        c = _distort(model, extra, u, v - ev); d = _distort(model, extra, u, v + ev)
        j00 = 1 + (b[0] - a[0]) / (2 * eu); j01 = (d[0] - c[0]) / (2 * ev)
        j10 = (b[1] - a[1]) / (2 * eu); j11 = 1 + (d[1] - c[1]) / (2 * ev)
        f0, f1 = u + du - u0, v + dv - v0
        det = j00 * j11 - j01 * j10
        su, sv = (j11 * f0 - j01 * f1) / det, (j00 * f1 - j10 * f0) / det
        u, v = u - su, v - sv
        if max(np.abs(su).max(initial=0.0), np.abs(sv).max(initial=0.0)) < 1e-12:
            break
    return np.stack([u, v], 1)


# --------------------------------------------------------------------------------------------
# Synthetic inputs -- seeded, dataset-free
# --------------------------------------------------------------------------------------------

from benchdata import polys_frame, frame_tensor  # noqa: E402,F401  (input generator shared with bench.py)


# --------------------------------------------------------------------------------------------
# "next" rows: recognition -> matching glue and projection refinement
# --------------------------------------------------------------------------------------------

def add_segmentations(logits: Tensor, filtering_threshold: float):
    """Reference localization/frame.py:96-121 for one frame ([N,C] logits): softmax, background pre-filter
    (kept only when at least 40 % of the keypoints survive), seg ids = argmax - 1.
    Returns (keep mask [N] over the ORIGINAL keypoints, seg_scores of the kept, seg_ids of the kept)."""
    scores = torch.softmax(logits, dim=-1)
    keep = torch.ones(logits.shape[0], dtype=torch.bool)
    if filtering_threshold > 0:
        non_bg = scores[:, 0] < filtering_threshold
        if non_bg.sum() >= 0.4 * scores.shape[0]:
            keep = non_bg
    return keep, scores[keep], logits[keep].max(dim=-1)[1] - 1


def process_segmentations(segs: Tensor, topk: int = 10):
    """Reference localization/multimap3d.py:348-379: [(sid, keypoint ids, score), ...]."""
    vals, ids = torch.topk(segs, k=segs.shape[-1], largest=True, dim=-1)
    vals, ids = vals.numpy(), ids.numpy()
    out, used = [], []
    for k in range(segs.shape[-1]):
        vk, ik = vals[:, k], ids[:, k]
        cand = []
        for sid in np.unique(ik):
            if sid == 0 or sid in used:
                continue
            used.append(sid)
            sel = np.where(ik == sid)[0]
            cand.append((sel.shape[0], sid, sel, np.mean(vk[sel])))
        for c in sorted(cand, key=lambda item: item[0], reverse=True):
            out.append((c[1], c[2], c[3]))
            if len(out) >= topk:
                return out
    return out


def match_by_projection(q_kpts: np.ndarray, q_descs: np.ndarray, xyz: np.ndarray, descs: np.ndarray, R: np.ndarray,
                        t: np.ndarray, K: np.ndarray, width: int, height: int, threshold: float):
    """Reference localization/singlemap3d.py:405-440: project map points, mask by visibility, descriptor
    distance sqrt(2-2qd+1e-6) (+100 outside a 2*threshold window), top-2, ratio 0.995.
    Returns (matched keypoint ids, matched map-point indices into ``xyz``)."""
    Xc = xyz @ R.T + t
    proj = (K @ Xc.T)
    u, v, z = proj[0] / proj[2], proj[1] / proj[2], proj[2]
    mask = (z > 0) & (z < 100) & (u >= 0) & (u < width) & (v >= 0) & (v < height)
    idx = np.nonzero(mask)[0]
    uv = np.stack([u[mask], v[mask]], 0)
    err = torch.sqrt(((torch.from_numpy(q_kpts)[..., None] - torch.from_numpy(uv)[None]) ** 2).sum(1))
    oor = err >= 2 * threshold
    dd = torch.sqrt(2 - 2 * torch.from_numpy(q_descs).float() @ torch.from_numpy(descs[mask]).float().t() + 1e-6)
    dd[oor] = dd[oor] + 100
    d, ids = torch.topk(dd, k=2, largest=False, dim=1)
    ok = ((d[:, 0] / d[:, 1]) <= 0.995) & (d[:, 0] < 100)
    ok = ok.numpy()
    return np.where(ok)[0], idx[ids.numpy()[ok, 0]], d.numpy()


# ------------------------------------------------------------------------------------------------
# NearestNeighbor matcher (reference localization/matchers/nearest_neighbor.py:5-56)
# ------------------------------------------------------------------------------------------------

def nn_find_nn(sim: Tensor, ratio_thresh, distance_thresh):
    """reference nearest_neighbor.py:5-16."""
    sim_nn, ind_nn = sim.topk(2 if ratio_thresh else 1, dim=-1, largest=True)
    dist_nn = 2 * (1 - sim_nn)
    mask = torch.ones(ind_nn.shape[:-1], dtype=torch.bool)
    if ratio_thresh:
        mask = mask & (dist_nn[..., 0] <= (ratio_thresh ** 2) * dist_nn[..., 1])
    if distance_thresh:
        mask = mask & (dist_nn[..., 0] <= distance_thresh ** 2)
    matches = torch.where(mask, ind_nn[..., 0], ind_nn.new_tensor(-1))
    scores = torch.where(mask, (sim_nn[..., 0] + 1) / 2, sim_nn.new_tensor(0))
    return matches, scores


def nearest_neighbor_forward(desc0: Tensor, desc1: Tensor, ratio_threshold=None, distance_threshold=None,
                             do_mutual_check: bool = True):
    """desc0 [B,D,N], desc1 [B,D,M] -> {'matches0', 'matching_scores0'}; reference nearest_neighbor.py:36-56 and
    mutual_check :19-24."""
    sim = torch.einsum('bdn,bdm->bnm', desc0, desc1)
    m0, s0 = nn_find_nn(sim, ratio_threshold, distance_threshold)
    if do_mutual_check:
        m1, _ = nn_find_nn(sim.transpose(1, 2), ratio_threshold, distance_threshold)
        inds0 = torch.arange(m0.shape[-1])
        loop = torch.gather(m1, -1, torch.where(m0 > -1, m0, m0.new_tensor(0)))
        ok = (m0 > -1) & (inds0 == loop)
        m0 = torch.where(ok, m0, m0.new_tensor(-1))
    return {'matches0': m0, 'matching_scores0': s0, 'sim': sim}


def select_with_mask_loops(keypoints, scores, descriptors, mask, topK=-1):
    """Literal restatement of the mask branch of extract_sfd2_return (reference nets/sfd2.py:502-571, with its
    ``np.float`` spelled ``float``): per-keypoint Python loops, kept as the checker of the vectorised product code."""
    labels, others = [], []
    kw, sw, dw, ko, so, do = [], [], [], [], [], []
    id_img = np.int32(mask[:, :, 2]) * 256 * 256 + np.int32(mask[:, :, 1]) * 256 + np.int32(mask[:, :, 0])
    for i in range(keypoints.shape[0]):
        x, y = keypoints[i, 0], keypoints[i, 1]
        gid = id_img[int(y), int(x)]
        if gid == 0:
            ko.append(keypoints[i]); so.append(scores[i]); do.append(descriptors[i]); others.append(0)
        else:
            kw.append(keypoints[i]); sw.append(scores[i]); dw.append(descriptors[i]); labels.append(gid)
    if topK > 0:
        if topK <= len(kw):
            idxes = np.array(sw, float).argsort()[::-1][:topK]
            keypoints = np.array(kw, float)[idxes]
            scores = np.array(sw, float)[idxes]
            labels = np.array(labels, np.int32)[idxes]
            descriptors = np.array(dw, float)[idxes]
        elif topK >= len(kw) + len(ko):
            keypoints, scores, descriptors = kw, sw, dw
            for i in range(len(others)):
                keypoints.append(ko[i]); scores.append(so[i]); descriptors.append(do[i]); labels.append(others[i])
        else:
            n = topK - len(kw)
            idxes = np.array(so, float).argsort()[::-1][:n]
            keypoints, scores, descriptors = kw, sw, dw
            for i in idxes:
                keypoints.append(ko[i]); scores.append(so[i]); descriptors.append(do[i]); labels.append(others[i])
    return {'keypoints': np.array(keypoints, float), 'descriptors': np.array(descriptors, float),
            'scores': np.array(scores, float), 'labels': np.array(labels, np.int32)}
