"""GPU parity of the HBM-bound feature kernels (K5-K9) against the CPU oracle / golden vectors.
Integer outputs (NMS map support, keypoint indices, order) are bit-exact; floats carry a stated
tolerance."""
import numpy as np
import pytest
import torch

from oracle import pram_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops(lib, dev):
    from pram_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize('hc,wc,layout', [(15, 20, 'nchw'), (60, 80, 'nhwc'), (7, 33, 'nhwc'), (1, 1, 'nchw')])
def test_score_map_vs_oracle(ops, dev, hc, wc, layout):
    g = torch.Generator().manual_seed(hc * 100 + wc)
    logits = torch.randn(2, 65, hc, wc, generator=g) * 4
    ref = O.score_map_from_logits(logits)
    x = logits.to(dev)
    if layout == 'nhwc':
        x = x.permute(0, 2, 3, 1).contiguous()  # physical NHWC
    else:
        x = x.permute(0, 2, 3, 1)  # NHWC view over NCHW memory (strided channels)
    out = ops.score_map(x)
    assert out.shape == ref.shape
    # tolerance: expf vs the CPU's vectorised exp, a few ulp of values <= 1
    assert torch.allclose(out.cpu(), ref, rtol=2e-6, atol=1e-9)


def test_score_map_resize_vs_oracle(ops, dev):
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(1, 65, 9, 12, generator=g)
    ref = O.score_map_from_logits(logits, 70, 93)  # not a multiple of 8 -> bilinear resize branch
    out = ops.score_map(logits.to(dev).permute(0, 2, 3, 1), 70, 93)
    assert torch.allclose(out.cpu(), ref, rtol=1e-5, atol=1e-8)


def _detect(ops, score, conf_th, min_kp, max_kp, border, radius=4, **kw):
    k, s, n, _, nms = ops.detect_keypoints(score, conf_th, min_kp, max_kp, border, radius=radius, return_nms=True, **kw)
    torch.cuda.synchronize()
    return k.cpu(), s.cpu(), n.cpu(), nms.cpu()


def test_nms_and_selection_vs_golden(ops, dev, golden):
    g = golden('sfd2_160x120.npz')
    score = torch.from_numpy(g['score_map']).to(dev)
    for r, key in ((4, 'nms4'), (3, 'nms3')):
        k, s, n, nms = _detect(ops, score, 0.005, 32, 64, 4, radius=r)
        assert np.array_equal(nms.numpy(), g[key]), f'NMS map (radius {r}) must be bit exact'
    k, s, n, nms = _detect(ops, score, 0.005, int(g['min_keypoints']), int(g['max_keypoints']), 4)
    nk = int(n[0])
    assert nk == g['keypoints'].shape[0]
    assert np.array_equal(k[0, :nk].numpy(), g['keypoints']), 'top-k keypoints / order must be bit exact'
    assert np.array_equal(s[0, :nk].numpy(), g['scores'])
    k, s, n, nms = _detect(ops, score, 0.005, int(g['min_keypoints']), 4096, 4)
    nk = int(n[0])
    assert np.array_equal(k[0, :nk].numpy(), g['keypoints_all']), 'row-major regime must be bit exact'
    assert np.array_equal(s[0, :nk].numpy(), g['scores_all'])


@pytest.mark.parametrize('h,w,seed', [(480, 640, 0), (97, 131, 1), (33, 65, 2), (8, 8, 3), (200, 64, 4)])
def test_nms_selection_random_maps(ops, dev, h, w, seed):
    """Ragged sizes (tile tails), batch of 3, quantised scores (exact ties / plateaus exercise the
    equality tests of simple_nms), both threshold regimes."""
    g = torch.Generator().manual_seed(seed)
    score = torch.rand(3, h, w, generator=g) ** 4 * 0.3
    score[1] = torch.round(score[1] * 64) / 64  # plateaus and exact ties
    score[2] *= 0.02  # few points above the threshold -> fallback to th/2 on this frame only
    nms_ref = O.simple_nms(score, 4)
    K = 100
    k, s, n, nms = _detect(ops, score.to(dev), 0.005, 20, K, 4, cap=h * w)
    assert torch.equal(nms, nms_ref)
    for b in range(3):
        kr, sr = O.select_keypoints(nms_ref[b], 0.005, 20, K, 4)
        nb = int(n[b])
        assert nb == kr.shape[0]
        if nb == 0:
            continue
        if torch.equal(k[b, :nb], kr):
            assert torch.equal(s[b, :nb], sr)
            continue
        # torch.topk's order among exactly equal scores is unspecified: compare up to ties
        assert torch.equal(s[b, :nb], sr), 'selected score multiset must match'
        ours = {(float(x), float(y)) for x, y in k[b, :nb]}
        theirs = {(float(x), float(y)) for x, y in kr}
        smin = float(sr.min())
        diff = ours ^ theirs
        for x, y in diff:
            assert float(nms_ref[b, int(y), int(x)]) == smin, 'only k-th-score ties may differ'


def test_selection_empty_and_overflow(ops, dev):
    score = torch.full((1, 64, 64), 1e-4)
    k, s, n, nms = _detect(ops, score.to(dev), 0.005, 0, 50, 4)
    assert int(n[0]) == 0
    # constant map: every pixel is a local maximum -> candidate overflow is reported, not silently wrong
    score = torch.full((1, 64, 64), 0.5)
    k, s, n, cnt = [t.cpu() if torch.is_tensor(t) else t for t in ops.detect_keypoints(score.to(dev), 0.005, 0, 50, 4, cap=256)]
    assert int(cnt[0]) == 64 * 64 > 256


@pytest.mark.parametrize('c,norm', [(128, True), (256, False)])
def test_sample_vs_oracle(ops, dev, c, norm):
    g = torch.Generator().manual_seed(c)
    h, w = 30, 40
    fmap = torch.randn(1, c, h, w, generator=g)
    kpts = torch.stack([torch.randint(0, w * 4, (300,), generator=g), torch.randint(0, h * 4, (300,), generator=g)], 1).float()
    ref = O.sample_map(kpts, fmap, 4, norm)  # [C,n]
    out = ops.sample_features(fmap.permute(0, 2, 3, 1).contiguous().to(dev), kpts[None].to(dev), None, 4, norm)
    # tolerance: 4-tap fp32 interpolation, different FMA contraction than the CPU kernel
    assert torch.allclose(out[0].t().cpu(), ref, rtol=1e-5, atol=2e-6)


def test_sample_golden(ops, dev, golden):
    g = golden('sfd2_160x120.npz')
    # scores are an exact gather
    sc = ops.gather_scores(torch.from_numpy(g['score_map']).to(dev), torch.from_numpy(g['keypoints'])[None].to(dev), None)
    assert np.array_equal(sc[0].cpu().numpy(), g['sample_scores'])


def test_posenc_vs_oracle(ops, dev):
    g = torch.Generator().manual_seed(0)
    k = torch.rand(2, 50, 2, generator=g) * torch.tensor([640., 480.])
    wr = torch.randn(32, 2, generator=g)
    enc = O.fourier_encoding(wr, O.normalize_keypoints(k, (1, 3, 480, 640)))  # [2,B,1,N,64]
    cos, sin = ops.posenc(k.to(dev), 640., 480., wr.to(dev))
    assert torch.allclose(cos.cpu().view(2, 50, 32), enc[0, :, 0, :, ::2], atol=2e-6)
    assert torch.allclose(sin.cpu().view(2, 50, 32), enc[1, :, 0, :, ::2], atol=2e-6)
