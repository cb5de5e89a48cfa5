"""CPU tests of host-side logic that mirrors reference code paths outside the kernels."""
import numpy as np
import pytest

from oracle import pram_oracle as O


@pytest.mark.parametrize('topK', [-1, 5, 40, 70, 1000])
def test_select_with_mask_matches_reference_loops(topK):
    """Vectorised mask-guided keypoint selection (export path) vs the literal loop restatement of
    reference nets/sfd2.py:502-571, for every topK regime (fewer / more than the labelled ones, everything, off)."""
    from pram_b200.nets.sfd2 import select_with_mask
    rs = np.random.RandomState(topK + 7)
    h, w, n = 60, 80, 120
    mask = np.zeros((h, w, 3), np.uint8)
    mask[10:30, 5:40] = (3, 0, 0)
    mask[35:55, 30:70] = (1, 2, 0)   # id = 1 + 512
    kp = np.stack([rs.uniform(0, w - 1e-3, n), rs.uniform(0, h - 1e-3, n)], 1)
    sc = rs.rand(n)
    de = rs.randn(n, 8)
    ref = O.select_with_mask_loops(kp.copy(), sc.copy(), de.copy(), mask, topK)
    out = select_with_mask(kp, sc, de, mask, topK)
    assert set(out) == set(ref)
    for k in ref:
        assert out[k].dtype == ref[k].dtype and np.array_equal(out[k], ref[k]), k
    assert 0 < (out['labels'] != 0).sum() < n


def test_match_features_batch_host_logic_cpu():
    """match_pairs groups pairs of equal shape into one batched matcher call, keeps AdaGML at batch 1, packs
    descriptors the way each matcher expects ([B,N,D] for the attentional ones, [B,D,N] for NN) and returns the
    reference's record dtypes -- checked on CPU with a recording stand-in for the matcher."""
    import torch
    from pram_b200.localization import match_features_batch as MFB

    class Recorder(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.calls = []

        def forward(self, data):
            self.calls.append({k: tuple(v.shape) for k, v in data.items() if torch.is_tensor(v)})
            b, n = data['keypoints0'].shape[:2]
            m0 = torch.arange(n)[None].repeat(b, 1) % data['keypoints1'].shape[1]
            return {'matches0': m0, 'matching_scores0': torch.full((b, n), 0.5)}

    def feat(n, seed, size=(640, 480)):
        r = np.random.RandomState(seed)
        return {'keypoints': r.rand(n, 2).astype(np.float32), 'descriptors': r.randn(128, n).astype(np.float32),
                'scores': r.rand(n).astype(np.float32), 'image_size': np.array(size)}
    store = {'q0': feat(50, 0), 'q1': feat(50, 1), 'q2': feat(70, 2), 'd0': feat(60, 3), 'd1': feat(60, 4)}
    pairs = [('q0', 'd0'), ('q1', 'd1'), ('q2', 'd0'), ('q0', 'd1')]
    rec = Recorder()
    out = MFB.match_pairs(MFB.confs['gml'], pairs, store, model=rec, device='cpu')
    assert len(rec.calls) == 2                      # three 50x60 pairs in one call, the 70x60 pair alone
    big = max(rec.calls, key=lambda c: c['keypoints0'][0])
    assert big['keypoints0'] == (3, 50, 2) and big['descriptors0'] == (3, 50, 128) and big['descriptors1'] == (3, 60, 128)
    assert big['image0'] == (1, 1, 480, 640)        # (h, w) from image_size = (w, h), like FeaturePairsDataset
    assert set(out) == {MFB.names_to_pair(*p) for p in pairs}
    r0 = out[MFB.names_to_pair('q2', 'd0')]
    assert r0['matches0'].dtype == np.int16 and r0['matches0'].shape == (70,) and r0['matching_scores0'].dtype == np.float16
    # NN layout and AdaGML batching rule
    rec = Recorder()
    MFB.match_pairs(MFB.confs['NNM'], pairs[:2], store, model=rec, device='cpu')
    assert rec.calls[0]['descriptors0'] == (2, 128, 50)
    rec = Recorder()
    MFB.match_pairs(MFB.confs['adagml'], pairs[:2], store, model=rec, device='cpu')
    assert len(rec.calls) == 2 and rec.calls[0]['keypoints0'][0] == 1
    assert MFB.names_to_pair('a/b.png', 'c/d.png') == 'a-b.png/c-d.png'


def test_find_2d_3d_matches_cpu():
    """find_2D_3D_matches (reference localization/pose_estimator.py:88-134) with a stand-in matcher on CPU: id remap
    through the valid-3D mask, observation threshold, +0.5 pixel-centre shift, ascending query order."""
    import torch
    from types import SimpleNamespace
    from pram_b200.localization.pose_estimator import find_2D_3D_matches
    rs = np.random.RandomState(3)
    n_q, n_db = 40, 30
    q = {'keypoints': rs.rand(n_q, 2).astype(np.float32) * 100, 'scores': rs.rand(n_q).astype(np.float32),
         'descriptors': rs.randn(n_q, 16).astype(np.float32), 'image_size': np.array([640, 480])}
    store = {'db.png': {'keypoints': rs.rand(n_db, 2).astype(np.float32), 'scores': rs.rand(n_db).astype(np.float32),
                        'descriptors': rs.randn(16, n_db).astype(np.float32), 'image_size': np.array([640, 480])}}
    ids3d = np.where(np.arange(n_db) % 3 == 0, -1, np.arange(n_db) + 500)
    valid = np.nonzero(ids3d != -1)[0]
    db_images = {0: SimpleNamespace(name='db.png', point3D_ids=ids3d)}
    points3D = {int(i): SimpleNamespace(xyz=rs.randn(3), image_ids=[0] * (int(i) % 3 + 1)) for i in ids3d if i != -1}

    class Stub(torch.nn.Module):  # query i -> (i mod n_valid)-th VALID database keypoint unless i % 5 == 0 (unmatched)
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1))

        def forward(self, data):
            assert data['keypoints1'].shape[1] == valid.size and data['descriptors1'].shape == (1, valid.size, 16)
            m = torch.full((1, n_q), -1, dtype=torch.long)
            idx = torch.tensor([i for i in range(n_q) if i % 5])
            m[0, idx] = idx % valid.size
            return {'matches0': m}
    mp3d, mkpq, mp3d_ids, q_ids = find_2D_3D_matches(q, 0, points3D, store, db_images, Stub(), obs_th=2)
    exp_q, exp_ids = [], []
    for i in range(n_q):
        if i % 5 == 0:
            continue
        pid = int(ids3d[valid[i % valid.size]])
        if len(points3D[pid].image_ids) >= 2:
            exp_q.append(i); exp_ids.append(pid)
    assert q_ids == exp_q and mp3d_ids == exp_ids and len(exp_q) > 3
    assert np.allclose(mkpq, q['keypoints'][exp_q].astype(float) + 0.5)
    assert np.allclose(mp3d, np.array([points3D[i].xyz for i in exp_ids]))


def test_singlemap3d_localize_with_ref_frame_cpu():
    """SingleMap3D.localize_with_ref_frame / .match (reference localization/singlemap3d.py:127-226) with stand-ins
    for the matcher and the pose operator on CPU: result keys and values, semantic (per-landmark) reference subsets,
    the (1, 3, W, H) image_shape quirk, the +0.5 shift handed to the pose operator, the failure convention, and that a
    reference frame is uploaded once."""
    import torch
    from types import SimpleNamespace
    from pram_b200.localization.singlemap3d import RefFrame, SingleMap3D
    rs = np.random.RandomState(0)
    cam = SimpleNamespace(width=640, height=480, model='PINHOLE', params=[525, 525, 320, 240])
    n_ref = 50
    segs = rs.randint(1, 4, n_ref)
    ref = RefFrame(cam, 7, np.hstack([rs.rand(n_ref, 2) * 400, rs.rand(n_ref, 1)]).astype(np.float32),
                   rs.randn(n_ref, 16).astype(np.float32), rs.randn(n_ref, 3), np.arange(n_ref) + 1000, segs, device='cpu')
    seen = []

    class Stub(torch.nn.Module):  # query i -> reference (i mod n1) for odd i
        def forward(self, data):
            seen.append((data['image_shape0'], data['image_shape1'], data['keypoints1'].shape[1]))
            n0, n1 = data['keypoints0'].shape[1], data['keypoints1'].shape[1]
            m = torch.full((1, n0), -1, dtype=torch.long)
            idx = torch.arange(1, n0, 2)
            m[0, idx] = idx % n1
            return {'matches0': m}
    calls = {}

    def pose_fn(p2d, p3d, camera, estimation_options=None, refinement_options=None):
        calls['args'] = (p2d.copy(), p3d.copy(), estimation_options)
        if calls.get('fail'):
            return None
        return {'cam_from_world': SimpleNamespace(rotation=SimpleNamespace(quat=np.array([0.1, 0.2, 0.3, 0.9])),
                                                  translation=np.array([1.0, 2.0, 3.0])),
                'num_inliers': 5, 'inliers': np.ones(p2d.shape[0], bool)}
    smap = SingleMap3D({'localization': {'threshold': 8}}, Stub(), {7: ref}, {2: [7, 9], 0: [7]},
                       {int(i): int(s) for i, s in zip(ref.point3D_ids, segs)}, device='cpu', pose_fn=pose_fn)
    nq = 30
    q = SimpleNamespace(camera=cam, descriptors=rs.randn(nq, 16).astype(np.float32),
                        keypoints=np.hstack([rs.rand(nq, 2) * 400, rs.rand(nq, 1)]).astype(np.float32))
    ids = np.arange(4, 24)
    ret = smap.localize_with_ref_frame(q, ids, sid=2, semantic_matching=True)
    sub = np.nonzero(segs == 2)[0]
    local = np.arange(ids.size)
    matched = local[1::2]
    ref_idx = sub[matched % sub.size]
    assert seen[-1] == ((1, 3, 640, 480), (1, 3, 640, 480), sub.size)
    assert ret['success'] and np.allclose(ret['qvec'], [0.9, 0.1, 0.2, 0.3]) and np.allclose(ret['tvec'], [1, 2, 3])
    assert np.array_equal(ret['matched_keypoint_ids'], ids[matched])
    assert np.allclose(ret['matched_keypoints'], q.keypoints[ids[matched], :2])
    assert np.allclose(ret['matched_xyzs'], ref.xyzs[ref_idx]) and np.array_equal(ret['matched_point3D_ids'], ref.point3D_ids[ref_idx])
    assert (ret['matched_sids'] == 2).all() and ret['reference_frame_id'] == 7
    assert np.allclose(ret['matched_ref_keypoints'], ref.keypoints[ref_idx, :2])
    assert np.allclose(calls['args'][0], q.keypoints[ids[matched], :2] + 0.5) and calls['args'][2] == {'ransac': {'max_error': 8}}
    # non-semantic path uses every reference keypoint; second call re-uses the resident copy
    smap.localize_with_ref_frame(q, ids, sid=2, semantic_matching=False)
    smap.localize_with_ref_frame(q, ids, sid=0, semantic_matching=True)   # sid 0 -> all keypoints too
    assert seen[-1][2] == n_ref and set(ref._dev) == {2, 'all'}
    # failure convention
    calls['fail'] = True
    bad = smap.localize_with_ref_frame(q, ids, sid=2)
    assert bad['success'] is False and bad['num_inliers'] == 0 and bad['inliers'].shape == (matched.size,) and not bad['inliers'].any()
    # match()
    m = smap.match({'keypoints': q.keypoints[:, :2], 'descriptors': q.descriptors, 'scores': q.keypoints[:, 2], 'camera': cam},
                   ref.get_keypoints())
    exp = np.arange(1, nq, 2)
    assert np.array_equal(m['matched_keypoint_ids'], exp) and np.allclose(m['matched_xyzs'], ref.xyzs[exp % n_ref])
    assert np.array_equal(m['matched_point3D_ids'], ref.point3D_ids[exp % n_ref])


def test_tracker_track_last_frame_cpu():
    """Tracker.track_last_frame (reference localization/tracker.py:162-233) with stand-in operators: matches without a
    3-D point (id -1) are dropped, ids / xyz / sids gathered in query order, +0.5 shift, result keys."""
    import torch
    from types import SimpleNamespace
    from pram_b200.localization.tracker import Tracker
    rs = np.random.RandomState(4)
    cam = SimpleNamespace(width=640, height=480)
    n0, n1 = 30, 25
    curr = SimpleNamespace(camera=cam, keypoints=rs.rand(n0, 3).astype(np.float32) * 100, descriptors=rs.randn(n0, 8).astype(np.float32))
    pids = np.where(np.arange(n1) % 4 == 0, -1, np.arange(n1) + 300)
    last = SimpleNamespace(camera=cam, keypoints=rs.rand(n1, 3).astype(np.float32) * 100, descriptors=rs.randn(n1, 8).astype(np.float32),
                           xyzs=rs.randn(n1, 3), point3D_ids=pids, seg_ids=rs.randint(0, 5, n1), reference_frame_id=11,
                           matched_scene_name='scene')

    class Stub(torch.nn.Module):
        def forward(self, data):
            m = torch.full((1, n0), -1, dtype=torch.long)
            idx = torch.arange(0, n0, 3)
            m[0, idx] = (idx * 7) % n1
            return {'matches0': m}
    got = {}

    def pose_fn(p2d, p3d, camera, estimation_options=None, refinement_options=None):
        got['p2d'], got['opt'] = p2d.copy(), estimation_options
        return {'cam_from_world': SimpleNamespace(rotation=SimpleNamespace(quat=np.array([0., 0., 0., 1.])), translation=np.zeros(3)),
                'num_inliers': p2d.shape[0], 'inliers': np.ones(p2d.shape[0], bool)}
    ret = Tracker(None, Stub(), {'localization': {'threshold': 12}}, device='cpu', pose_fn=pose_fn).track_last_frame(curr, last)
    q = np.arange(0, n0, 3); r = (q * 7) % n1
    ok = pids[r] >= 0
    q, r = q[ok], r[ok]
    assert np.array_equal(ret['matched_keypoint_ids'], q) and np.array_equal(ret['matched_point3D_ids'], pids[r])
    assert np.allclose(ret['matched_xyzs'], last.xyzs[r]) and np.array_equal(ret['matched_sids'], last.seg_ids[r])
    assert np.allclose(ret['matched_keypoints'], curr.keypoints[q, :2]) and np.allclose(ret['matched_ref_keypoints'], last.keypoints[r, :2])
    assert np.allclose(got['p2d'], curr.keypoints[q, :2] + 0.5) and got['opt'] == {'ransac': {'max_error': 12}}
    assert ret['success'] and np.allclose(ret['qvec'], [1, 0, 0, 0]) and ret['reference_frame_id'] == 11 and ret['matched_scene_name'] == 'scene'


def test_bench_workloads_and_step_bound():
    """bench.py metadata helpers: the SURVEY.md 8d closed forms (250.8 / 677 / 1911 GFLOP per frame) and the
    step-level tensor ceiling; the workload table matches the BASELINE.json shapes."""
    import json
    import bench
    assert abs(bench.algorithmic_gflop_per_frame(480, 640, 1024, 113) - 250.8) < 0.1
    assert abs(bench.algorithmic_gflop_per_frame(768, 1024, 2048, 161) - 677.1) < 0.5
    assert abs(bench.algorithmic_gflop_per_frame(1200, 1600, 4096, 513) - 1911.0) < 2.0
    b = bench.whole_step_bound(1000.0, 2, 'bf16x3')
    json.dumps(b)
    assert b['tensor_bound_frames_per_s'] > 2000 and 0 < b['frac_of_tensor_bound'] < 1
    assert bench.WORKLOADS['7scenes']['kpts'] == 1024 and bench.WORKLOADS['cambridge']['h'] == 768 and bench.WORKLOADS['aachen']['kpts'] == 4096
    saved = (bench.H, bench.W, bench.KPTS, bench.NCLASS, bench.FOCAL, bench.MAX_ERROR, bench.METRIC)
    try:
        assert bench.set_workload('cambridge', 0) == 16 and (bench.H, bench.W, bench.KPTS, bench.NCLASS) == (768, 1024, 2048, 161)
        assert '1024x768' in bench.METRIC and bench.set_workload('7scenes', 4) == 4
        assert bench.METRIC == 'localization frames/sec (640x480, 1024 kpts)'
    finally:
        bench.H, bench.W, bench.KPTS, bench.NCLASS, bench.FOCAL, bench.MAX_ERROR, bench.METRIC = saved


def test_committed_profile_summaries_parse():
    """bench.py reads its roofline traffic figure from profiles/*.json; every committed summary must be valid JSON (an empty
    file once took the multi-GPU bench down) and the conv3b capture must carry the DRAM byte count."""
    import json
    from pathlib import Path
    prof = Path(__file__).resolve().parents[1] / 'profiles'
    files = sorted(prof.glob('r02*.json'))   # (two round-1 captures carry an NCCL banner line in front of the JSON)
    assert files
    for f in files:
        json.loads(f.read_text())
    d = json.loads((prof / 'r02b_conv3b_ncu_full.json').read_text())
    assert d['dram_bytes'] > 1e9 and 'gemm_tc_kernel' in d['kernel']


def test_attention_launch_merging_rule():
    """nets/_blocks._mergeable: both token sets of a GML block share one attention launch only when the two segments are
    adjacent, equally shaped, no mean-attention output is requested, and padded batches bring the concatenated counts."""
    import torch
    from pram_b200.nets import _blocks as BL
    seg = ((0, 4, 100), (400, 4, 100))
    assert BL._mergeable(seg, None, None)
    assert not BL._mergeable(seg, [torch.zeros(1)], None)                      # AdaGML asks for the per-key attention mass
    assert not BL._mergeable(((0, 4, 100), (400, 4, 90)), None, None)           # M != N
    assert not BL._mergeable(((0, 4, 100), (512, 4, 100)), None, None)          # not adjacent
    assert not BL._mergeable(((0, 4, 100),), None, None)                        # SegNetViT: one segment
    c = torch.full((4,), 100, dtype=torch.int32)
    assert not BL._mergeable(seg, None, [c, c])                                 # counts without the concatenated form
    assert BL._mergeable(seg, None, [c, None, torch.cat([c, c])])
    saved = BL.MERGE_SETS
    try:
        BL.MERGE_SETS = False
        assert not BL._mergeable(seg, None, None)
    finally:
        BL.MERGE_SETS = saved
