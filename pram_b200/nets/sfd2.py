"""B200-native SFD2 feature operator -- drop-in for reference ``nets/sfd2.py``.

Same class name, constructor, state-dict key schema (``load_state_dict(strict=True)`` of the shipped
checkpoint works) and method contracts as the reference:

* ``ResNet4x.extract_local_global(data, config)``  reference nets/sfd2.py:269-346
* ``ResNet4x.sample(score_map, semi_descs, kpts, s, norm_desc)``  nets/sfd2.py:348-369
* ``ResNet4x.det(x)`` / ``ResNet4x.forward(batch)``  nets/sfd2.py:172-233
* ``extract_sfd2_return(model, img, ...)``  nets/sfd2.py:386-589
* ``load_sfd2(weight_path)``  nets/sfd2.py:592-596

The ``nn.Module`` containers below exist only to own the parameters under the reference's names; no
torch operator runs in the forward path.  Inference repacks the weights once (BatchNorm folded into
the preceding convolution, NHWC / tap-major layouts) and then calls hand-written sm_100a kernels
through the C ABI (``include/pram_b200.h``).  Activations are NHWC on the device; the NCHW tensors the
reference API promises are returned as channels-last *views* (same shape, same values).
There is no CPU path: tensors must be CUDA tensors.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from .. import _lib, ops

RGB_mean = [0.485, 0.456, 0.406]
RGB_std = [0.229, 0.224, 0.225]


def _unit(cin: int, cout: int, stride: int = 1) -> nn.Sequential:
    # parameter holder for conv3x3 + BN (+ReLU); indices 0/1 give the reference's key names
    return nn.Sequential(nn.Conv2d(cin, cout, 3, stride, 1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class ResBlock(nn.Module):
    """Parameter holder with the reference's key names (nets/sfd2.py:94-105)."""

    def __init__(self, inplanes: int, outplanes: int, groups: int = 32):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, outplanes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(outplanes)
        self.conv2 = nn.Conv2d(outplanes, outplanes, 3, 1, 1, groups=groups, bias=False)
        self.bn2 = nn.BatchNorm2d(outplanes)
        self.conv3 = nn.Conv2d(outplanes, outplanes, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(outplanes)


def _fold(conv_w: torch.Tensor, conv_b: Optional[torch.Tensor], bn: Optional[nn.BatchNorm2d]):
    """Fold eval-mode BatchNorm into the convolution (float64 on the host, once)."""
    w = conv_w.detach().double()
    b = conv_b.detach().double() if conv_b is not None else torch.zeros(w.shape[0], dtype=torch.float64,
                                                                         device=w.device)
    if bn is not None:
        g = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
        w = w * g.view(-1, 1, 1, 1)
        b = (b - bn.running_mean.detach().double()) * g + bn.bias.detach().double()
    return w, b


def _tapmajor(w: torch.Tensor) -> torch.Tensor:
    """[Cout,Cin,kh,kw] -> [kh*kw, Cin, Cout] fp32 contiguous."""
    co, ci, kh, kw = w.shape
    return w.permute(2, 3, 1, 0).reshape(kh * kw, ci, co).float().contiguous()


class ResNet4x(nn.Module):
    default_config = {
        'conf_th': 0.005,
        'remove_borders': 4,
        'min_keypoints': 128,
        'max_keypoints': 4096,
    }

    def __init__(self, inputdim: int = 3, outdim: int = 128, desc_compressor=None):
        super().__init__()
        self.outdim = outdim
        self.desc_compressor = desc_compressor
        self.conv1a = _unit(inputdim, 64)
        self.conv1b = _unit(64, 64, 2)
        self.conv2a = _unit(64, 128)
        self.conv2b = _unit(128, 128, 2)
        self.conv3a = _unit(128, 256)
        self.conv3b = _unit(256, 256)
        self.conv4 = nn.Sequential(ResBlock(256, 256), ResBlock(256, 256), ResBlock(256, 256))
        self.convPa = nn.Sequential(nn.Conv2d(256, 256, 3, 2, 1), nn.BatchNorm2d(256), nn.ReLU(inplace=True),
                                    nn.Conv2d(256, 256, 3, 1, 1))
        self.convDa = nn.Sequential(nn.Conv2d(256, 256, 3, 1, 1), nn.BatchNorm2d(256), nn.ReLU(inplace=True),
                                    nn.Conv2d(256, 256, 3, 1, 1))
        self.convPb = nn.Conv2d(256, 65, 1)
        self.convDb = nn.Conv2d(256, outdim, 1)
        self._packed: Optional[Dict[str, torch.Tensor]] = None
        # 'bf16x3' : tcgen05 tensor cores, error-compensated split-bf16 operands (default; ~fp32 results)
        # 'bf16'   : tcgen05 tensor cores, plain bf16 operands (fastest; keypoint order not stable)
        # 'fp32'   : CUDA-core fp32 kernels (exact-arithmetic reference path)
        self.precision = 'bf16x3'
        # precision of the DESCRIPTOR head only (convDa.0, convDa.3, convDb; 46.6 of the 133 GFLOP of the stack): None = same
        # as the trunk; 'f16' = single-pass IEEE fp16 operands (11-bit mantissas, fp32 accumulation) -- descriptors are
        # sampled, L2-normalised and matched with a tolerance, they never feed the discontinuous keypoint selection
        self.desc_precision = None
        self.eval()

    @property
    def compute_dtype(self) -> str:
        base = {'bf16x3': 'bf16x3 (split-bf16 tcgen05, fp32 accumulate)', 'bf16': 'bf16', 'fp32': 'f32'}[self.precision]
        return base + (' + fp16 descriptor head' if (self.desc_precision == 'f16' and self.precision != 'fp32') else '')

    def set_precision(self, precision: str, desc_precision: Optional[str] = None):
        if precision not in ('bf16x3', 'bf16', 'fp32') or desc_precision not in (None, 'f16'):
            raise ValueError((precision, desc_precision))
        self.precision, self.desc_precision = precision, desc_precision
        return self

    # -- weight repacking ---------------------------------------------------------------------
    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    def prepare(self) -> Dict[str, torch.Tensor]:
        """One-time repack: BN folded (fp64), tap-major [taps,Cin,Cout] fp32; grouped conv as
        [tap][ci][co][group]."""
        if self._packed is not None:
            return self._packed
        dev = self.convPb.weight.device
        if dev.type != 'cuda':
            raise _lib.PramError('ResNet4x must be moved to a CUDA device (.cuda()) before inference: '
                                 'pram_b200 has no CPU path')
        pk: Dict[str, torch.Tensor] = {}

        def put(name, w, b):
            pk[name + '.w'] = _tapmajor(w).to(dev)
            pk[name + '.b'] = b.float().contiguous().to(dev)
            # tensor-core layout: [taps, Cout, Cin] (K-major rows), split into bf16 hi/lo planes
            co, ci, kh, kw = w.shape
            if ci % 8 == 0:
                wt = w.permute(2, 3, 0, 1).reshape(kh * kw, co, ci).float().contiguous().to(dev)
                pk[name + '.tc'] = ops.split_bf16(wt)
                if name in ('convDa.0', 'convDa.3', 'convDb'):
                    pk[name + '.tc16'] = ops.as_f16_plane(wt)

        for name in ('conv1a', 'conv1b', 'conv2a', 'conv2b', 'conv3a', 'conv3b'):
            seq = getattr(self, name)
            put(name, *_fold(seq[0].weight, seq[0].bias, seq[1]))
        for i, blk in enumerate(self.conv4):
            put(f'conv4.{i}.c1', *_fold(blk.conv1.weight, None, blk.bn1))
            w2, b2 = _fold(blk.conv2.weight, None, blk.bn2)  # [256, 8, 3, 3]
            g = 32
            w2 = w2.view(g, 8, 8, 3, 3)  # [group, co, ci, kh, kw]
            pk[f'conv4.{i}.c2.w'] = w2.permute(3, 4, 2, 1, 0).reshape(9, 8, 8, g).float().contiguous().to(dev)
            pk[f'conv4.{i}.c2.b'] = b2.float().contiguous().to(dev)
            put(f'conv4.{i}.c3', *_fold(blk.conv3.weight, None, blk.bn3))
        for head in ('convPa', 'convDa'):
            seq = getattr(self, head)
            put(head + '.0', *_fold(seq[0].weight, seq[0].bias, seq[1]))
            put(head + '.3', *_fold(seq[3].weight, seq[3].bias, None))
        put('convPb', *_fold(self.convPb.weight, self.convPb.bias, None))
        put('convDb', *_fold(self.convDb.weight, self.convDb.bias, None))
        self._packed = pk
        return pk

    # -- the conv stack -----------------------------------------------------------------------
    def _trunk(self, image: torch.Tensor) -> Dict[str, torch.Tensor]:
        """image [B,3,H,W] (normalised) -> NHWC fp32 maps.  K1-K4 of SURVEY.md section 2b."""
        _lib.require_cuda(image, 'image')
        pk = self.prepare()
        if self.precision != 'fp32':
            return self._trunk_tc(image, pk, 3 if self.precision == 'bf16x3' else 1)
        x = image.float().permute(0, 2, 3, 1).contiguous()  # NHWC, 3 channels (layout plumbing only)
        c = ops.conv_f32
        o1a = c(x, pk['conv1a.w'], pk['conv1a.b'], 3, 1, True)
        o1b = c(o1a, pk['conv1b.w'], pk['conv1b.b'], 3, 2, True)
        o2a = c(o1b, pk['conv2a.w'], pk['conv2a.b'], 3, 1, True)
        o2b = c(o2a, pk['conv2b.w'], pk['conv2b.b'], 3, 2, True)
        o3a = c(o2b, pk['conv3a.w'], pk['conv3a.b'], 3, 1, True)
        o3b = c(o3a, pk['conv3b.w'], pk['conv3b.b'], 3, 1, True)
        o4 = o3b
        for i in range(3):
            t = c(o4, pk[f'conv4.{i}.c1.w'], pk[f'conv4.{i}.c1.b'], 1, 1, True)
            t = ops.gconv3x3_f32(t, pk[f'conv4.{i}.c2.w'], pk[f'conv4.{i}.c2.b'], True)
            o4 = c(t, pk[f'conv4.{i}.c3.w'], pk[f'conv4.{i}.c3.b'], 1, 1, True, res=o4)
        p = c(o4, pk['convPa.0.w'], pk['convPa.0.b'], 3, 2, True)
        p = c(p, pk['convPa.3.w'], pk['convPa.3.b'], 3, 1, False)
        logits = c(p, pk['convPb.w'], pk['convPb.b'], 1, 1, False)
        d = c(o4, pk['convDa.0.w'], pk['convDa.0.b'], 3, 1, True)
        d = c(d, pk['convDa.3.w'], pk['convDa.3.b'], 3, 1, False)
        desc = c(d, pk['convDb.w'], pk['convDb.b'], 1, 1, False)
        ops.l2norm_rows_(desc, desc.shape[-1])
        return {'out1b': o1b, 'out2b': o2b, 'out3b': o3b, 'out4': o4, 'logits': logits, 'desc': desc}

    def _trunk_tc(self, image: torch.Tensor, pk, split: int) -> Dict[str, torch.Tensor]:
        """Tensor-core conv stack: tcgen05 implicit GEMMs fed by TMA, activations as split-bf16 NHWC planes
        (phase-split in front of the three stride-2 convolutions); CUDA cores only for conv1a (Cin=3,
        HBM-bound); the 32-group 3x3 convolutions run on warp-level bf16 MMAs (gconv_mma.cu)."""
        b, _, h, w = image.shape
        ct = ops.conv_tc
        T = lambda n: pk[n + '.tc']
        ps1a, _ = ops.conv1a(image, pk['conv1a.w'], pk['conv1a.b'], split)
        h2, w2 = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        o1b = ct(ps1a, T('conv1b'), pk['conv1b.b'], 3, 2, True, split, out_shape_hw=(h2, w2))['bf']
        ps2a = ct(o1b, T('conv2a'), pk['conv2a.b'], 3, 1, True, split, want_bf=False, want_ps=True)['ps']
        h4, w4 = (h2 - 1) // 2 + 1, (w2 - 1) // 2 + 1
        o2b = ct(ps2a, T('conv2b'), pk['conv2b.b'], 3, 2, True, split, out_shape_hw=(h4, w4))['bf']
        o3a = ct(o2b, T('conv3a'), pk['conv3a.b'], 3, 1, True, split)['bf']
        # the ResBlock residuals are taken from the producer's split-bf16 planes (bf16x3 mode: hi + lo carries ~16 mantissa
        # bits): the 256-channel maps at 1/4 resolution are written as fp32 only once, for out4 (mid_features of the API)
        planes_res = split == 3
        o3 = ct(o3a, T('conv3b'), pk['conv3b.b'], 3, 1, True, split, want_f32=not planes_res)
        cur_bf, cur_f32 = o3['bf'], o3.get('f32')
        o3b_bf = cur_bf
        last = None
        for i in range(3):
            t = ct(cur_bf, T(f'conv4.{i}.c1'), pk[f'conv4.{i}.c1.b'], 1, 1, True, split)['bf']
            t = ops.gconv3x3_tc(t, pk[f'conv4.{i}.c2.w'], pk[f'conv4.{i}.c2.b'], True, split)
            # the last block feeds convPa (phase-split planes), the API's mid_features (fp32) and the descriptor head: as
            # ONE fp16 plane written by the same epilogue when that head runs single-pass fp16 (no separate cast kernel,
            # and no hi / lo planes that nobody reads)
            f16_tail = (i == 2) and planes_res and self.desc_precision == 'f16'
            last = ct(t, T(f'conv4.{i}.c3'), pk[f'conv4.{i}.c3.b'], 1, 1, True, split, res=None if planes_res else cur_f32,
                      res_bf=cur_bf if planes_res else None, want_f32=(i == 2) or not planes_res, want_ps=(i == 2),
                      want_bf=not f16_tail, want_h16=f16_tail)
            cur_bf, cur_f32 = last.get('bf'), last.get('f32')
        h8, w8 = (h4 - 1) // 2 + 1, (w4 - 1) // 2 + 1
        p = ct(last['ps'], T('convPa.0'), pk['convPa.0.b'], 3, 2, True, split, out_shape_hw=(h8, w8))['bf']
        p = ct(p, T('convPa.3'), pk['convPa.3.b'], 3, 1, False, split)['bf']
        logits = ct(p, T('convPb'), pk['convPb.b'], 1, 1, False, split, want_f32=True, want_bf=False)['f32']
        if self.desc_precision == 'f16':
            T16 = lambda n: pk[n + '.tc16']
            d = last['h16'] if 'h16' in last else ops.as_f16_plane(cur_f32)   # out4 as one fp16 plane
            d = ct(d, T16('convDa.0'), pk['convDa.0.b'], 3, 1, True, 1, f16=True)['bf']
            d = ct(d, T16('convDa.3'), pk['convDa.3.b'], 3, 1, False, 1, f16=True)['bf']
            desc = ct(d, T16('convDb'), pk['convDb.b'], 1, 1, False, 1, want_f32=True, want_bf=False, l2norm=True, f16=True)['f32']
        else:
            d = ct(cur_bf, T('convDa.0'), pk['convDa.0.b'], 3, 1, True, split)['bf']
            d = ct(d, T('convDa.3'), pk['convDa.3.b'], 3, 1, False, split)['bf']
            desc = ct(d, T('convDb'), pk['convDb.b'], 1, 1, False, split, want_f32=True, want_bf=False, l2norm=True)['f32']
        return {'out1b': o1b, 'out2b': o2b, 'out3b': o3b_bf, 'out4': cur_f32, 'logits': logits, 'desc': desc}

    @staticmethod
    def _nchw(x_nhwc) -> torch.Tensor:
        if isinstance(x_nhwc, ops.Split):  # tensor-core path keeps some maps only as split-bf16 planes
            x_nhwc = x_nhwc.float()
        return x_nhwc.permute(0, 3, 1, 2)  # channels-last view with the reference's NCHW shape

    # -- reference API ------------------------------------------------------------------------
    @torch.no_grad()
    def det(self, x: torch.Tensor):
        t = self._trunk(x)
        return ops.score_map(t['logits']), self._nchw(t['desc'])

    @torch.no_grad()
    def forward(self, batch: dict) -> dict:
        t = self._trunk(batch['image'])
        logits = self._nchw(t['logits'])
        semi = torch.softmax(logits, dim=1)[:, :-1]  # training-time convenience output only
        return {'dense_features': self._nchw(t['desc']), 'scores': ops.score_map(t['logits']),
                'logits': logits, 'semi_map': semi}

    extract_patches = forward

    @torch.no_grad()
    def extract_local_global(self, data: dict, config: Optional[dict] = None) -> dict:
        cfg = {**self.default_config, **(config or {})}
        out = self.extract_batched(data['image'], cfg)
        n = out['num_keypoints'].tolist()  # the only host sync of the feature stage
        b = len(n)
        if any(c > out['cand_cap'] for c in out['cand_count'].tolist()):
            # plateau image: more NMS survivors than the candidate buffer -- redo with a full buffer
            out = self.extract_batched(data['image'], cfg, cap=data['image'].shape[-1] * data['image'].shape[-2])
            n = out['num_keypoints'].tolist()
        if cfg['max_keypoints'] < 0 and any(v > out['keypoints'].shape[1] for v in out['num_valid'].tolist()):
            raise _lib.PramError(f"max_keypoints = -1 (unlimited) found {max(out['num_valid'].tolist())} keypoints, more than "
                                 f"the selection kernel's {out['keypoints'].shape[1]} slots; pass max_keypoints <= 4096")
        return {
            'score_map': out['score_map'],
            'desc_map': self._nchw(out['desc_map_nhwc']),
            'mid_features': self._nchw(out['mid_features_nhwc']),
            'global_descriptors': [self._nchw(v) for v in out['global_nhwc']],
            'keypoints': [out['keypoints'][i, :n[i]] for i in range(b)],
            'scores': tuple(out['scores'][i, :n[i]] for i in range(b)),
            'descriptors': [out['descriptors'][i, :n[i]].t() for i in range(b)],  # [128, n] views
        }

    @torch.no_grad()
    def extract_batched(self, image: torch.Tensor, cfg: Optional[dict] = None, cap: Optional[int] = None) -> dict:
        """Device-resident, sync-free form of ``extract_local_global`` (padded [B,K,...] outputs +
        per-frame counts); the frame-parallel runner consumes this directly."""
        cfg = {**self.default_config, **(cfg or {})}
        b, _, ih, iw = image.shape
        t = self._trunk(image)
        score = ops.score_map(t['logits'], ih, iw)
        if cap is None:
            cap = ops.default_cand_cap(ih, iw, 4)
        kpts, scs, n, cand_count, n_valid = ops.detect_keypoints(score, cfg['conf_th'], cfg['min_keypoints'],
                                                                 cfg['max_keypoints'], cfg['remove_borders'], radius=4,
                                                                 cap=cap, return_valid=True)
        desc = ops.sample_features(t['desc'], kpts, n, 4, True)
        return {'score_map': score, 'desc_map_nhwc': t['desc'], 'mid_features_nhwc': t['out4'],
                'global_nhwc': [t['out1b'], t['out2b'], t['out3b'], t['out4']],
                'keypoints': kpts, 'scores': scs, 'descriptors': desc, 'num_keypoints': n,
                'cand_count': cand_count, 'cand_cap': cap, 'num_valid': n_valid, 'logits': t['logits']}

    @torch.no_grad()
    def sample(self, score_map: torch.Tensor, semi_descs: torch.Tensor, kpts: torch.Tensor, s: int = 4,
               norm_desc: bool = True):
        """(scores [n], descriptors [C,n]); reference nets/sfd2.py:348-369."""
        _lib.require_cuda(semi_descs, 'semi_descs')
        fm = semi_descs.permute(0, 2, 3, 1)
        if not fm.is_contiguous() or fm.dtype != torch.float32:
            fm = fm.float().contiguous()
        k = kpts.reshape(1, -1, 2).float().contiguous()
        d = ops.sample_features(fm[:1] if fm.shape[0] != 1 else fm, k, None, s, norm_desc)
        sc = ops.gather_scores(score_map[:1], k, None)
        return sc[0], d[0].t()


class DescriptorCompressor(nn.Module):
    """Parameter holder for the optional 1x1 Conv1d compressor (reference nets/sfd2.py:372-383); not on
    the hot path (``desc_compressor=None`` everywhere the reference constructs ResNet4x)."""

    def __init__(self, inputdim: int, outdim: int):
        super().__init__()
        self.inputdim, self.outdim = inputdim, outdim
        self.conv = nn.Conv1d(inputdim, outdim, 1)


def extract_sfd2_return(model: ResNet4x, img: torch.Tensor, conf_th: float = 0.001, mask=None, topK: int = -1,
                        min_keypoints: int = 0, **kwargs):
    """Offline-export variant (reference nets/sfd2.py:386-589): NMS radius 3, strict ``>`` threshold,
    score-descending order, border 4, sampling with x/(w/2)-1, float64 numpy outputs.
    ``img``: [1,3,H,W] (or [3,H,W]) RGB in [0,1], not normalised.  ``mask`` ([H,W,3] uint8 label image, BGR-packed
    ids) selects labelled keypoints first (reference nets/sfd2.py:502-571; host-side numpy like the reference)."""
    dev = next(model.parameters()).device
    x = img.reshape(1, 3, img.shape[-2], img.shape[-1]).to(dev).float()
    mean = torch.tensor(RGB_mean, device=dev).view(1, 3, 1, 1)
    std = torch.tensor(RGB_std, device=dev).view(1, 3, 1, 1)
    x = (x - mean) / std
    H, W = x.shape[2:]
    scales = kwargs.get('scales', [1.0])
    all_pts, all_desc = [], []
    for s in scales:
        xi = x if s == 1.0 else torch.nn.functional.interpolate(x, size=(int(H * s), int(W * s)), mode='bilinear',
                                                                align_corners=True)
        nh, nw = xi.shape[2:]
        t = model._trunk(xi)
        heat = ops.score_map(t['logits'], nh, nw)
        # every candidate above the threshold, best first (K = 4096 is the kernel's and the
        # reference config's ceiling, extract_features.py:73)
        # border test against the ORIGINAL width / height even on a rescaled map, as the reference does (:447-451)
        kpts, scs, n, cc = ops.detect_keypoints(heat, conf_th, 0, 4096, 4, radius=3, strict=True, fallback=False,
                                                border_hi=(H - 4, W - 4))
        if int(cc[0]) > ops.default_cand_cap(nh, nw, 3):   # plateau image: redo with a full candidate buffer
            kpts, scs, n, cc = ops.detect_keypoints(heat, conf_th, 0, 4096, 4, radius=3, strict=True, fallback=False,
                                                    border_hi=(H - 4, W - 4), cap=nh * nw)
        n0 = int(n[0])
        if n0 == 0:
            continue
        k = kpts[:, :n0]
        order = torch.argsort(scs[0, :n0], descending=True, stable=True)  # row-major when n <= K
        k, sc = k[:, order], scs[0, :n0][order]
        # reference samples with g = x/(w/2) - 1 on the desc map: express through pixel coords of the
        # generic sampler: (g+1)/2*(wd-1) with wd = desc width  ->  handled by s=1 on a coordinate remap
        dh, dw = t['desc'].shape[1:3]
        g = torch.stack([k[0, :, 0] / (float(nw) / 2.) - 1., k[0, :, 1] / (float(nh) / 2.) - 1.], -1)
        pix = torch.stack([(g[:, 0] + 1) / 2 * (dw - 1), (g[:, 1] + 1) / 2 * (dh - 1)], -1)
        d = _sample_at_map_coords(t['desc'], pix)
        pts = torch.cat([k[0] * torch.tensor([W / nw, H / nh], device=dev), sc[:, None]], 1)
        all_pts.append(pts.double().cpu().numpy())
        all_desc.append(d.double().cpu().numpy())
    if not all_pts:
        return None, None, None
    pts = np.vstack(all_pts)
    desc = np.vstack(all_desc)
    keypoints, scores = pts[:, :2], pts[:, 2]
    if mask is not None:
        return select_with_mask(keypoints, scores, desc, mask, topK)
    if topK > 0:
        idx = np.array(scores, dtype=float).argsort()[::-1][:topK]
        keypoints, scores, desc = keypoints[idx], scores[idx], desc[idx]
    return {'keypoints': np.array(keypoints, dtype=float), 'descriptors': np.array(desc, dtype=float),
            'scores': np.array(scores, dtype=float)}


def select_with_mask(keypoints: np.ndarray, scores: np.ndarray, descriptors: np.ndarray, mask: np.ndarray, topK: int = -1):
    """Mask-guided selection of the export path (reference nets/sfd2.py:502-571), vectorised: keypoints on a labelled
    pixel (id = B + 256 G + 65536 R of ``mask[int(y), int(x)]`` != 0) come first; ``topK`` keeps the best labelled
    ones, or all labelled ones plus the best unlabelled ones.  Same outputs and orders as the reference's loops
    (including its ``topK <= 0`` quirk: every keypoint is returned, ``labels`` only covers the labelled ones)."""
    id_img = np.int32(mask[:, :, 2]) * 256 * 256 + np.int32(mask[:, :, 1]) * 256 + np.int32(mask[:, :, 0])
    gid = id_img[keypoints[:, 1].astype(int), keypoints[:, 0].astype(int)]
    w_idx, o_idx = np.nonzero(gid != 0)[0], np.nonzero(gid == 0)[0]
    labels = gid[w_idx]
    if topK > 0:
        if topK <= w_idx.size:
            sel = w_idx[np.array(scores[w_idx], float).argsort()[::-1][:topK]]
            labels = gid[sel]
        elif topK >= w_idx.size + o_idx.size:
            sel = np.concatenate([w_idx, o_idx])
            labels = np.concatenate([labels, np.zeros(o_idx.size, labels.dtype)])
        else:
            extra = o_idx[np.array(scores[o_idx], float).argsort()[::-1][:topK - w_idx.size]]
            sel = np.concatenate([w_idx, extra])
            labels = np.concatenate([labels, np.zeros(extra.size, labels.dtype)])
        keypoints, scores, descriptors = keypoints[sel], scores[sel], descriptors[sel]
    return {'keypoints': np.array(keypoints, float), 'descriptors': np.array(descriptors, float),
            'scores': np.array(scores, float), 'labels': np.array(labels, np.int32)}


def _sample_at_map_coords(desc_nhwc: torch.Tensor, pix: torch.Tensor) -> torch.Tensor:
    """Bilinear sample at map-pixel coordinates, then L2 normalise -> [n, C].  The generic sampler maps
    keypoint k to map coordinate ((k - off)/div*2-1+1)/2*(w-1); with s chosen so that off = 0 and
    div = w-1... the export path's normalisation is not of that form, so the coordinates are
    pre-inverted here: k' = pix * div/(w-1) + off."""
    b, h, w, c = desc_nhwc.shape
    s = 4
    off = s / 2 - 0.5
    divx, divy = w * s - s / 2 - 0.5, h * s - s / 2 - 0.5
    k = torch.stack([pix[:, 0] * (divx / (w - 1)) + off, pix[:, 1] * (divy / (h - 1)) + off], -1)
    return ops.sample_features(desc_nhwc[:1], k[None].contiguous(), None, s, True)[0]


def load_sfd2(weight_path: str) -> ResNet4x:
    net = ResNet4x(inputdim=3, outdim=128)
    net.load_state_dict(torch.load(weight_path, map_location='cpu', weights_only=False)['state_dict'], strict=True)
    return net
