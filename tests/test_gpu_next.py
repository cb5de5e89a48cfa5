"""'Next' rows of SURVEY.md section 8f on the device: Frame.add_segmentations, process_segmentations,
projection-based matching (K18) -- against the CPU restatements of the reference code."""
import numpy as np
import pytest
import torch

from oracle import pram_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops(lib, dev):
    from pram_b200 import ops as _ops
    return _ops


def test_add_segmentations(ops, dev):
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(700, 113, generator=g) * 3
    logits[:300, 0] += 9  # confident background on part of the keypoints
    keep, scores, ids = O.add_segmentations(logits, 0.95)
    bg, sid, nb, probs = ops.segmentation(logits.to(dev), 0.95, want_probs=True)
    assert torch.allclose(probs.cpu(), torch.softmax(logits, -1), atol=1e-6)
    assert torch.equal(nb.cpu(), torch.softmax(logits, -1)[:, 0] < 0.95)
    assert nb.sum().item() >= 0.4 * 700 and torch.equal(nb.cpu(), keep)
    assert torch.equal(sid.cpu().long()[keep], ids)


@pytest.mark.parametrize('topk,n,c', [(20, 1024, 113), (30, 300, 161), (10, 2048, 513)])
def test_process_segmentations(ops, dev, topk, n, c):
    g = torch.Generator().manual_seed(topk)
    # a few dominant landmarks + noise, so that several ranks are needed to collect topk entries
    logits = torch.randn(n, c, generator=g)
    dom = torch.randint(1, min(c, 12), (n,), generator=g)
    logits[torch.arange(n), dom] += 4
    ref = O.process_segmentations(logits, topk)
    out = ops.rank_landmarks(logits[None].to(dev), None, topk, max_ranks=16)
    ne = int(out['n'][0])
    assert ne == len(ref)
    lab = out['label_at_rank'][0].cpu()
    for e, (sid, ids, score) in enumerate(ref):
        assert int(out['sid'][0, e]) == int(sid)
        k = int(out['rank'][0, e])
        assert np.array_equal(torch.nonzero(lab[k] == int(sid))[:, 0].numpy(), ids)
        assert int(out['count'][0, e]) == len(ids)
        assert abs(float(out['score'][0, e]) - float(score)) < 1e-5


def test_match_by_projection(ops, dev):
    rs = np.random.RandomState(0)
    m, n, th = 600, 5000, 8.0
    f, w, h = 525.0, 640, 480
    K = np.array([[f, 0, w / 2], [0, f, h / 2], [0, 0, 1.0]])
    ang = 0.1
    R = O.quat_to_rotmat(np.array([np.cos(ang / 2), 0, np.sin(ang / 2), 0.0]))
    t = np.array([0.05, -0.02, 0.1])
    # map points: the first m project (with sub-pixel noise) onto the query keypoints, the rest are clutter
    kp = np.stack([rs.uniform(8, w - 8, m), rs.uniform(8, h - 8, m)], 1)
    z = rs.uniform(1, 6, m)
    Xc = np.stack([(kp[:, 0] - w / 2) / f * z, (kp[:, 1] - h / 2) / f * z, z], 1)
    X_in = (Xc - t) @ R
    X_cl = rs.uniform(-6, 6, (n - m, 3)) + np.array([0, 0, 4.0])
    xyz = np.concatenate([X_in, X_cl]).astype(np.float32).astype(np.float64)
    qd = rs.randn(m, 128); qd /= np.linalg.norm(qd, axis=1, keepdims=True)
    dd = rs.randn(n, 128); dd /= np.linalg.norm(dd, axis=1, keepdims=True)
    dd[:m] = qd + 0.05 * rs.randn(m, 128); dd[:m] /= np.linalg.norm(dd[:m], axis=1, keepdims=True)
    qd, dd = qd.astype(np.float32), dd.astype(np.float32)
    kpn = (kp + rs.normal(0, 0.7, kp.shape)).astype(np.float32)
    ids_q, ids_m, d_ref = O.match_by_projection(kpn.astype(np.float64), qd, xyz, dd, R, t, K, w, h, th)
    match, d0, d1 = ops.match_by_projection(torch.from_numpy(kpn).to(dev), torch.from_numpy(qd).to(dev),
                                            torch.from_numpy(xyz).float().to(dev), torch.from_numpy(dd).to(dev), R, t, f, f,
                                            w / 2, h / 2, w, h, th)
    match = match.cpu().numpy()
    assert np.allclose(d0.cpu().numpy(), d_ref[:, 0], atol=2e-3) and np.allclose(d1.cpu().numpy(), d_ref[:, 1], atol=2e-3)
    ref_full = -np.ones(m, np.int64)
    ref_full[ids_q] = ids_m
    decisive = np.abs(d_ref[:, 0] / d_ref[:, 1] - 0.995) > 2e-3
    assert np.array_equal(match[decisive], ref_full[decisive])
    assert (match[:m] == np.arange(m)).mean() > 0.9  # the planted correspondences are found
