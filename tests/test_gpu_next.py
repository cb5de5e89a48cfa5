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


@pytest.mark.parametrize('conf', [{}, {'ratio_threshold': 0.9}, {'distance_threshold': 0.7, 'do_mutual_check': False},
                                  {'ratio_threshold': 0.95, 'distance_threshold': 0.9}])
def test_nearest_neighbor_matcher(dev, lib, conf):
    """NN matcher plugin (tcgen05 similarity + device top-2 / ratio / distance / mutual check) vs the CPU restatement of
    reference localization/matchers/nearest_neighbor.py.  Matches are compared exactly on decisive rows: the top-1 /
    top-2 similarity gap and the distance of each test to its threshold exceed the bf16x3 GEMM tolerance (1e-5)."""
    from pram_b200.localization import matchers
    from pram_b200.localization.base_model import dynamic_load
    Model = dynamic_load(matchers, 'nearest_neighbor')
    g = torch.Generator().manual_seed(5)
    d0 = torch.nn.functional.normalize(torch.randn(2, 128, 300, generator=g), dim=1)
    d1 = torch.nn.functional.normalize(torch.randn(2, 128, 257, generator=g), dim=1)
    d1[:, :, :100] = torch.nn.functional.normalize(d0[:, :, 50:150] + 0.1 * torch.randn(2, 128, 100, generator=g), dim=1)
    full = {**Model.default_conf, **conf}
    ref = O.nearest_neighbor_forward(d0, d1, **full)
    out = Model(conf).eval().to(dev)({'descriptors0': d0.to(dev), 'descriptors1': d1.to(dev)})
    m, s = out['matches0'].cpu(), out['matching_scores0'].cpu()
    tol = 1e-5
    sim = ref['sim']

    def decisive(sm):  # rows whose decisions cannot flip within tol
        top = sm.topk(3, dim=-1).values
        d = 2 * (1 - top)
        ok = (top[..., 0] - top[..., 1]) > tol
        if full['ratio_threshold']:
            ok &= (d[..., 0] - full['ratio_threshold'] ** 2 * d[..., 1]).abs() > 8 * tol
        if full['distance_threshold']:
            ok &= (d[..., 0] - full['distance_threshold'] ** 2).abs() > 8 * tol
        return ok
    dec0 = decisive(sim)
    if full['do_mutual_check']:
        dec1 = decisive(sim.transpose(1, 2))
        # a row is decisive if its own decision and the reverse decision of its candidate column are
        cand = sim.argmax(-1)
        dec0 &= torch.gather(dec1, 1, cand)
    assert dec0.float().mean() > 0.95
    assert torch.equal(m[dec0], ref['matches0'][dec0])
    assert torch.allclose(s[dec0], ref['matching_scores0'][dec0], atol=1e-5)
    if not conf:  # plain mutual NN finds the planted correspondences (thresholded variants may legitimately reject them)
        assert (m[:, 50:150] == torch.arange(100)).float().mean() > 0.9


def test_match_features_batch_and_find_2d_3d(dev, lib):
    """match_features_batch on in-memory feature stores (the reference's h5 layout: descriptors [D, N]) with pairs of
    equal size matched in one batched call, records in the reference's dtypes; find_2D_3D_matches on the same store
    (reference pose_estimator.py:88-134) against a direct restatement of its per-match loop."""
    from types import SimpleNamespace
    from pram_b200.localization import match_features_batch as MFB
    from pram_b200.localization.pose_estimator import find_2D_3D_matches
    rs = np.random.RandomState(0)

    def feat(n, seed):
        r = np.random.RandomState(seed)
        d = r.randn(128, n).astype(np.float32)
        d /= np.linalg.norm(d, axis=0, keepdims=True)
        return {'keypoints': r.uniform(0, 480, (n, 2)).astype(np.float32), 'descriptors': d,
                'scores': r.rand(n).astype(np.float32), 'image_size': np.array([640, 480])}
    store = {f'db/{i}.png': feat(200, i) for i in range(3)}
    q = feat(200, 100)
    perm = rs.permutation(200)
    q['descriptors'][:, :150] = store['db/0.png']['descriptors'][:, perm[:150]]  # 150 planted correspondences
    store['query/a.png'] = q
    pairs = [('query/a.png', f'db/{i}.png') for i in range(3)]
    recs = MFB.main(MFB.confs['NNM'], pairs, store)
    assert set(recs) == {MFB.names_to_pair(*p) for p in pairs}
    r0 = recs[MFB.names_to_pair(*pairs[0])]
    assert r0['matches0'].dtype == np.int16 and r0['matching_scores0'].dtype == np.float16 and r0['matches0'].shape == (200,)
    assert (r0['matches0'][:150] == perm[:150]).all()
    # batched call == one call per pair
    for p in pairs:
        single = MFB.match_pairs(MFB.confs['NNM'], [p], store)
        assert np.array_equal(single[MFB.names_to_pair(*p)]['matches0'], recs[MFB.names_to_pair(*p)]['matches0'])
    # find_2D_3D_matches
    ids3d = np.where(rs.rand(200) < 0.7, np.arange(200) + 1000, -1)
    db_images = {7: SimpleNamespace(name='db/0.png', point3D_ids=ids3d)}
    points3D = {int(i): SimpleNamespace(xyz=rs.randn(3), image_ids=list(range(int(i) % 4))) for i in ids3d if i != -1}
    nn_model = MFB.load_matcher(MFB.confs['NNM'], dev)

    class _RowMajorNN(torch.nn.Module):  # feature_matching packs descriptors [B, N, D] (the GML convention)
        def forward(self, data):
            return nn_model({'descriptors0': data['descriptors0'].transpose(1, 2).contiguous(),
                             'descriptors1': data['descriptors1'].transpose(1, 2).contiguous()})
    matcher = _RowMajorNN()
    q = dict(q, descriptors=np.ascontiguousarray(q['descriptors'].T))  # query features as the extractor hands them over: [N, D]
    mp3d, mkpq, mp3d_ids, q_ids = find_2D_3D_matches(q, 7, points3D, store, db_images, matcher, obs_th=2)
    from pram_b200.localization.pose_estimator import feature_matching
    dbf = store['db/0.png']
    m = feature_matching(q, {'keypoints': dbf['keypoints'], 'scores': dbf['scores'], 'descriptors': dbf['descriptors'].T,
                             'db_3D_ids': ids3d, 'image_size': dbf['image_size']}, matcher)
    exp_q, exp_ids = [], []
    for idx in range(m.shape[0]):  # the reference's loop
        if m[idx] == -1 or ids3d[m[idx]] == -1:
            continue
        if len(points3D[int(ids3d[m[idx]])].image_ids) < 2:
            continue
        exp_q.append(idx); exp_ids.append(int(ids3d[m[idx]]))
    assert q_ids == exp_q and mp3d_ids == exp_ids and len(q_ids) > 20
    assert np.allclose(mkpq, q['keypoints'][exp_q].astype(float) + 0.5)
    assert np.allclose(mp3d, np.array([points3D[i].xyz for i in exp_ids]))
