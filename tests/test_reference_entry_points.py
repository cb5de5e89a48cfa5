"""CPU pins of the host-side entry points against the REFERENCE'S OWN functions / methods (imported from /root/reference
with stub modules for pycolmap / h5py, ``ref_loc`` fixture): same stand-in matcher and pose operator on both sides, so any
difference is a difference in the control flow, id bookkeeping or tensor packing around the two device operators.
Build container only (skips where the reference tree is not mounted)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch


class CodeMatcher(torch.nn.Module):
    """Stand-in matcher: keypoint i of set 0 matches keypoint j of set 1 when their descriptors carry the same integer code
    in component 0 (codes <= 0 never match).  Works for [B,N,D] descriptors (attentional matchers' layout)."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))
        self.calls = []

    def forward(self, data):
        c0, c1 = data['descriptors0'][0, :, 0].round().long(), data['descriptors1'][0, :, 0].round().long()
        self.calls.append({k: (tuple(v.shape) if torch.is_tensor(v) else v) for k, v in data.items()})
        m = torch.full((1, c0.shape[0]), -1, dtype=torch.long)
        for i, c in enumerate(c0.tolist()):
            if c > 0:
                j = torch.nonzero(c1 == c)
                if j.numel():
                    m[0, i] = int(j[0, 0])
        return {'matches0': m}


def make_pose_fn(log, fail_below=0, inlier_rule=None):
    """Stand-in for pycolmap.absolute_pose_estimation: deterministic in its inputs, records them."""
    def pose_fn(p2d, p3d, camera, estimation_options=None, refinement_options=None):
        p2d, p3d = np.asarray(p2d, float), np.asarray(p3d, float)
        log.append((p2d.copy(), p3d.copy(), estimation_options))
        if p2d.shape[0] < fail_below:
            return None
        inl = (np.floor(p2d[:, 0]).astype(int) % 3 != 0) if inlier_rule is None else inlier_rule(p2d)
        q = np.array([0.1, 0.2, 0.3, 0.9]) + 1e-3 * p2d.shape[0]
        return {'cam_from_world': SimpleNamespace(rotation=SimpleNamespace(quat=q), translation=np.array([1.0, 2.0, p3d[:, 2].sum()])),
                'num_inliers': int(inl.sum()), 'inliers': inl}
    return pose_fn


def _scene(seed=0, n_db=4, n_q=60):
    """A tiny SfM scene: database images with features in an h5-like store, 3-D points with observation lists."""
    rs = np.random.RandomState(seed)
    n_pts = 80
    points3D = {1000 + k: SimpleNamespace(xyz=rs.randn(3) * 2 + [0, 0, 5], image_ids=[10 + (k + i) % n_db for i in range(k % 4 + 1)]) for k in range(n_pts)}
    store, db_images = {}, {}
    for d in range(n_db):
        n = 30 + 5 * d
        codes = rs.permutation(np.arange(1, n_pts + 1))[:n]
        desc = rs.randn(8, n).astype(np.float32)
        desc[0] = codes
        p3d_ids = np.where(rs.rand(n) < 0.25, -1, 1000 + codes - 1)
        name = f'db/{d}.png'
        store[name] = {'keypoints': rs.rand(n, 2).astype(np.float32) * 100, 'scores': rs.rand(n).astype(np.float32),
                       'descriptors': desc, 'image_size': np.array([640, 480])}
        db_images[10 + d] = SimpleNamespace(name=name, point3D_ids=p3d_ids, qvec=np.array([1.0, 0, 0, d]), tvec=np.array([d, 0, 0.0]))
    qcodes = rs.permutation(np.arange(1, n_pts + 1))[:n_q]
    qdesc = rs.randn(8, n_q).astype(np.float32)
    qdesc[0] = qcodes
    store['query.png'] = {'keypoints': (rs.rand(n_q, 2) * 100).astype(np.float32), 'scores': rs.rand(n_q).astype(np.float32),
                          'descriptors': qdesc, 'image_size': np.array([640, 480])}
    qinfo = ('SIMPLE_PINHOLE', 640, 480, [500.0, 320.0, 240.0])
    return store, db_images, points3D, qinfo


def _same(a, b, skip=('time',)):
    assert set(a) - set(skip) == set(b) - set(skip), (sorted(a), sorted(b))
    for k in a:
        if k in skip:
            continue
        va, vb = a[k], b[k]
        if isinstance(va, SimpleNamespace) or k == 'cam_from_world':
            continue
        if isinstance(va, (np.ndarray, list, tuple)) and not isinstance(va, str):
            va, vb = np.asarray(va), np.asarray(vb)
            assert va.shape == vb.shape and va.dtype.kind == vb.dtype.kind, k
            assert np.array_equal(va, vb), k
        else:
            assert va == vb, (k, va, vb)


def test_feature_matching_and_2d3d_vs_reference(ref_loc):
    from pram_b200.localization import pose_estimator as P
    store, db_images, points3D, _ = _scene(1)
    q = {k: (v.transpose() if k == 'descriptors' else v) for k, v in store['query.png'].items()}
    for obs_th in (0, 2, 3):
        for db_id in db_images:
            ref = ref_loc.pose_estimator.find_2D_3D_matches(q, db_id, points3D, store, db_images, CodeMatcher(), obs_th=obs_th)
            our = P.find_2D_3D_matches(q, db_id, points3D, store, db_images, CodeMatcher(), obs_th=obs_th)
            assert np.array_equal(ref[0], our[0]) and np.array_equal(ref[1], our[1]) and ref[2] == our[2] and ref[3] == our[3]
    db = store['db/0.png']
    for ids in (None, db_images[10].point3D_ids, np.full(db['keypoints'].shape[0], -1)):
        dbd = {'keypoints': db['keypoints'], 'scores': db['scores'], 'descriptors': db['descriptors'].transpose(),
               'image_size': db['image_size'], 'db_3D_ids': ids}
        m_ref, m_our = CodeMatcher(), CodeMatcher()
        r = ref_loc.pose_estimator.feature_matching(q, dbd, m_ref)
        o = P.feature_matching(q, dbd, m_our)
        assert np.array_equal(r, o)
        assert m_ref.calls == m_our.calls  # identical tensor packing (shapes, image0/1 from image_size)


def test_covisibility_frames_vs_reference(ref_loc):
    from pram_b200.localization import pose_estimator as P
    _, db_images, points3D, _ = _scene(2, n_db=6)
    for fid in db_images:
        for k in (2, 50):
            assert list(ref_loc.pose_estimator.get_covisibility_frames(fid, db_images, points3D, k)) == \
                   list(P.get_covisibility_frames(fid, db_images, points3D, k))


@pytest.mark.parametrize('fail_below', [0, 10 ** 6])
def test_pose_estimator_hloc_vs_reference(ref_loc, fail_below):
    from pram_b200.localization import pose_estimator as P
    store, db_images, points3D, qinfo = _scene(3)
    log_r, log_o = [], []
    ref_loc.pycolmap.absolute_pose_estimation = make_pose_fn(log_r, fail_below)
    ref_loc.pycolmap.Camera = lambda **kw: kw
    db_ids = list(db_images)
    r = ref_loc.pose_estimator.pose_estimator_hloc('query.png', qinfo, db_ids, db_images, points3D, store, 12, None, CodeMatcher(),
                                                   log_info='')
    o = P.pose_estimator_hloc('query.png', qinfo, db_ids, db_images, points3D, store, 12, None, CodeMatcher(), log_info='',
                              pose_fn=make_pose_fn(log_o, fail_below))
    _same(r, o)
    assert len(log_r) == len(log_o) == 1 and np.array_equal(log_r[0][0], log_o[0][0]) and np.array_equal(log_r[0][1], log_o[0][1])
    assert log_r[0][2] == log_o[0][2] == {'ransac': {'max_error': 12}}
    if fail_below == 0:
        assert r['num_inliers'] > 0 and len(r['points3D_ids']) == r['num_inliers']


@pytest.mark.parametrize('fail_below', [0, 10 ** 6])
def test_pose_refinement_vs_reference(ref_loc, fail_below):
    from pram_b200.localization import pose_estimator as P
    store, db_images, points3D, qinfo = _scene(4, n_db=5)
    q = {k: (v.transpose() if k == 'descriptors' else v) for k, v in store['query.png'].items()}
    cam = {'model': qinfo[0], 'width': qinfo[1], 'height': qinfo[2], 'params': qinfo[3]}
    log_r, log_o = [], []
    ref_loc.pycolmap.absolute_pose_estimation = make_pose_fn(log_r, fail_below)
    kw = dict(query_data=q, query_cam=cam, feature_file=store, db_frame_id=11, db_images=db_images, points3D=points3D,
              covisibility_frame=3, obs_th=2, opt_th=9, qvec=np.array([1.0, 0, 0, 0]), tvec=np.zeros(3), log_info='')
    r = ref_loc.pose_estimator.pose_refinement(matcher=CodeMatcher(), **kw)
    o = P.pose_refinement(matcher=CodeMatcher(), pose_fn=make_pose_fn(log_o, fail_below), **kw)
    _same(r, o)
    assert np.array_equal(log_r[0][0], log_o[0][0]) and log_o[0][2] == {'ransac': {'max_error': 9}}


def test_pose_estimator_iterative_vs_reference(ref_loc):
    """The paths of the reference that run: first candidate too few 3-D matches / too few inliers, a later one succeeds."""
    from pram_b200.localization import pose_estimator as P
    store, db_images, points3D, qinfo = _scene(5, n_db=5)
    ref_loc.pycolmap.Camera = lambda **kw: kw
    db_ids = list(db_images)
    for inlier_th in (1, 6, 9):
        log_r, log_o = [], []
        rule = lambda p2d: np.arange(p2d.shape[0]) < (p2d.shape[0] // 2 + int(p2d[:, 0].sum()) % 3)
        ref_loc.pycolmap.absolute_pose_estimation = make_pose_fn(log_r, 0, rule)
        r = ref_loc.pose_estimator.pose_estimator_iterative('query.png', qinfo, db_ids, db_images, points3D, store, 12, None,
                                                            CodeMatcher(), inlier_th=inlier_th, log_info='', obs_th=0)
        o = P.pose_estimator_iterative('query.png', qinfo, db_ids, db_images, points3D, store, 12, None, CodeMatcher(),
                                       inlier_th=inlier_th, log_info='', obs_th=0, pose_fn=make_pose_fn(log_o, 0, rule))
        assert len(log_r) == len(log_o) and r['order'] == o['order'] > 0
        _same(r, o)
    # nothing localises: the pose of the first database image, num_inliers -1 (the reference raises here, see docstring)
    o = P.pose_estimator_iterative('query.png', qinfo, db_ids, db_images, points3D, store, 12, None, CodeMatcher(), inlier_th=50,
                                   log_info='', pose_fn=make_pose_fn([], 10 ** 6))
    assert o['num_inliers'] == -1 and np.array_equal(o['qvec'], db_images[db_ids[0]].qvec)
    # covisibility refinement on the winning candidate (the reference raises TypeError on this path)
    o = P.pose_estimator_iterative('query.png', qinfo, db_ids, db_images, points3D, store, 12, None, CodeMatcher(), inlier_th=1,
                                   log_info='', do_covisibility_opt=True, covisibility_frame=3, pose_fn=make_pose_fn([], 0))
    assert o['num_inliers'] > 0 and len(o['points3D_ids']) == o['num_inliers']


# ---- SingleMap3D / Tracker against the reference's own methods -----------------------------------------------------------------

def _map_scene(ref_loc, seed=0):
    """The same small landmark map as reference objects (RefFrame / Point3D / SingleMap3D.__new__) and as ours."""
    from pram_b200.localization.singlemap3d import RefFrame, SingleMap3D
    rs = np.random.RandomState(seed)
    Point3D = __import__('localization.point3d', fromlist=['Point3D']).Point3D
    cam_model = SimpleNamespace(name='PINHOLE')
    cam = SimpleNamespace(id=1, model=cam_model, width=640, height=480, params=[500.0, 500.0, 320.0, 240.0])
    n_pts, frame_ids = 120, [3, 5, 8, 9]
    p3d = {}
    for k in range(n_pts):
        pid = 2000 + k
        code = k + 1
        desc = rs.randn(16).astype(np.float32) * 0.05
        desc[0] = code
        p3d[pid] = Point3D(id=pid, xyz=np.array([rs.uniform(-2, 2), rs.uniform(-1.5, 1.5), rs.uniform(4, 8)]), error=rs.rand(),
                           refframe_id=-1, seg_id=int(k % 4) + 1, descriptor=desc, frame_ids=[frame_ids[(k + i) % 4] for i in range(k % 3 + 1)])
    ref_frames_r, ref_frames_o = {}, {}
    for fid in frame_ids:
        ids = np.array([pid for pid, p in p3d.items() if fid in p.frame_ids])
        xyz = np.array([p3d[i].xyz for i in ids])
        desc = np.array([p3d[i].descriptor for i in ids])
        kp = np.hstack([rs.rand(ids.size, 2) * [640, 480], rs.rand(ids.size, 1)]).astype(np.float32)
        segs = np.array([p3d[i].seg_id for i in ids])
        rf = ref_loc.refframe.RefFrame(camera=cam, id=fid, qvec=np.array([1.0, 0, 0, 0]), tvec=np.zeros(3), point3D_ids=ids, keypoints=kp)
        rf.descriptors, rf.xyzs, rf.keypoint_segs = desc, xyz, segs
        ref_frames_r[fid] = rf
        ref_frames_o[fid] = RefFrame(cam, fid, kp, desc, xyz, ids, segs, device='cpu')
    seg_ref = {1: [3, 5], 2: [5], 3: [8], 4: [9], 0: [3]}
    config = {'localization': {'threshold': 8, 'covisibility_frame': 3}}
    r = ref_loc.singlemap3d.SingleMap3D.__new__(ref_loc.singlemap3d.SingleMap3D)
    r.config, r.point3Ds, r.reference_frames, r.seg_ref_frame_ids, r.start_sid = config, p3d, ref_frames_r, seg_ref, 0
    r.build_covisibility_graph(frame_ids=frame_ids, n_frame=3)
    o = SingleMap3D(config, None, ref_frames_o, seg_ref, {pid: p.seg_id for pid, p in p3d.items()}, device='cpu', point3Ds=p3d)
    return r, o, p3d, cam, rs


def _query_frame(p3d, cam, rs, n=70):
    """A query frame looking at the map with pose (I, 0): keypoints near the projections of some 3-D points."""
    pids = rs.permutation(list(p3d))[:n]
    kp = np.zeros((n, 3), np.float32)
    desc = np.zeros((n, 16), np.float32)
    for i, pid in enumerate(pids):
        X = p3d[pid].xyz
        kp[i, :2] = [500 * X[0] / X[2] + 320 + rs.uniform(-3, 3), 500 * X[1] / X[2] + 240 + rs.uniform(-3, 3)]
        kp[i, 2] = rs.rand()
        desc[i] = p3d[pid].descriptor + rs.randn(16).astype(np.float32) * 0.01
    desc /= np.linalg.norm(desc, axis=1, keepdims=True)
    fr = SimpleNamespace(camera=cam, keypoints=kp, descriptors=desc, qvec=np.array([1.0, 0, 0, 0]), tvec=np.zeros(3),
                         reference_frame_id=5, tracking_status=True, seg_ids=np.array([p3d[i].seg_id for i in pids]),
                         get_intrinsics=lambda: np.array([[500.0, 0, 320], [0, 500.0, 240], [0, 0, 1]]))
    return fr, pids


def test_singlemap3d_methods_vs_reference(ref_loc):
    r, o, p3d, cam, rs = _map_scene(ref_loc, 0)
    assert {k: list(v) for k, v in r.covisible_graph.items()} == {k: list(v) for k, v in o.covisible_graph.items()}
    fr, pids = _query_frame(p3d, cam, rs)
    fr.descriptors[:, 0] = [(p - 2000 + 1) for p in pids]   # integer codes for the stand-in matcher
    for rf in list(r.reference_frames.values()) + list(o.reference_frames.values()):
        rf.descriptors = rf.descriptors.copy()
        rf._dev = {} if hasattr(rf, '_dev') else None
    ids = np.arange(5, 60)
    for sid, semantic in ((1, True), (2, True), (2, False), (0, True)):
        log_r, log_o = [], []
        r.matcher, o.matcher = CodeMatcher(), CodeMatcher()
        ref_loc.pycolmap.absolute_pose_estimation = make_pose_fn(log_r)
        o.pose_fn = make_pose_fn(log_o)
        a = r.localize_with_ref_frame(fr, ids, sid=sid, semantic_matching=semantic)
        b = o.localize_with_ref_frame(fr, ids, sid=sid, semantic_matching=semantic)
        _same(a, b)
        assert r.matcher.calls == o.matcher.calls and np.array_equal(log_r[0][0], log_o[0][0]) and log_r[0][2] == log_o[0][2]
    ref_loc.pycolmap.absolute_pose_estimation = make_pose_fn([], fail_below=10 ** 6)
    o.pose_fn = make_pose_fn([], fail_below=10 ** 6)
    _same(r.localize_with_ref_frame(fr, ids, sid=3), o.localize_with_ref_frame(fr, ids, sid=3))
    q = {'keypoints': fr.keypoints[:, :2], 'descriptors': fr.descriptors, 'scores': fr.keypoints[:, 2], 'camera': cam}
    _same(r.match(q, r.reference_frames[8].get_keypoints()), o.match(q, o.reference_frames[8].get_keypoints()))
    some = list(pids[:30])
    assert list(r.find_reference_frames(some, r.covisible_graph.keys())) == list(o.find_reference_frames(some, o.covisible_graph.keys()))
    for sid in (1, 2, 3):
        assert r.check_semantic_consistency(fr, sid, 0.2) == o.check_semantic_consistency(fr, sid, 0.2)
    # refinement by matching over the covisible frames (min / max trials and confidence handed to the pose operator)
    fr.matched_keypoints, fr.matched_keypoint_ids = fr.keypoints[:4, :2], np.arange(4)
    fr.matched_point3D_ids = pids[:4]
    log_r, log_o = [], []
    r.matcher, o.matcher = CodeMatcher(), CodeMatcher()
    ref_loc.pycolmap.absolute_pose_estimation = make_pose_fn(log_r)
    o.pose_fn = make_pose_fn(log_o)
    _same(r.refine_pose(fr, 'matching'), o.refine_pose(fr, 'matching'))
    assert log_o[0][2] == log_r[0][2] == {'ransac': {'max_error': 8, 'min_num_trials': 1000, 'max_num_trials': 10000, 'confidence': 0.995}}


def test_refine_pose_by_projection_vs_reference(ref_loc, monkeypatch):
    """SingleMap3D.refine_pose_by_projection against the reference method (its inline torch code runs on CPU here).
    The device operator of our side (projection + similarity GEMM + masked top-2) is replaced by the oracle's
    ``match_by_projection`` -- which this test thereby pins to the reference's inline implementation
    (singlemap3d.py:405-440); the CUDA operator is pinned to the same oracle function in tests/test_gpu_next.py."""
    from oracle import pram_oracle as O
    from pram_b200 import ops
    r, o, p3d, cam, rs = _map_scene(ref_loc, 1)
    fr, pids = _query_frame(p3d, cam, rs)
    for p in p3d.values():   # unit descriptors, as the map stores them
        p.descriptor = p.descriptor / np.linalg.norm(p.descriptor)
    o._p3d_table = None

    def fake(q_kpts, q_descs, xyz, descs, R, t, fx, fy, cx, cy, width, height, threshold, ratio=0.995, split=3):
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
        kid, pidx, _ = O.match_by_projection(q_kpts.numpy().astype(np.float32), q_descs.numpy(), xyz.double().numpy(), descs.numpy(),
                                             np.asarray(R, float), np.asarray(t, float), K, width, height, threshold)
        m = torch.full((q_kpts.shape[0],), -1, dtype=torch.long)
        m[torch.from_numpy(kid)] = torch.from_numpy(pidx)
        return m, None, None
    monkeypatch.setattr(ops, 'match_by_projection', fake)
    log_r, log_o = [], []
    ref_loc.pycolmap.absolute_pose_estimation = make_pose_fn(log_r)
    o.pose_fn = make_pose_fn(log_o)
    a = r.refine_pose(fr, 'projection')
    b = o.refine_pose(fr, 'projection')
    assert a['matched_keypoint_ids'].size > 20
    _same(a, b)
    assert np.array_equal(log_r[0][0], log_o[0][0]) and np.array_equal(log_r[0][1], log_o[0][1]) and log_r[0][2] == log_o[0][2]


def test_tracker_vs_reference(ref_loc):
    from pram_b200.localization.tracker import Tracker
    rs = np.random.RandomState(4)
    cam = SimpleNamespace(width=640, height=480)
    n0, n1 = 40, 35

    def frame(n, with3d):
        d = rs.randn(n, 8).astype(np.float32)
        d[:, 0] = rs.permutation(50)[:n] + 1
        f = SimpleNamespace(camera=cam, keypoints=(rs.rand(n, 3) * 100).astype(np.float32), descriptors=d)
        if with3d:
            f.xyzs, f.seg_ids = rs.randn(n, 3), rs.randint(0, 5, n)
            f.point3D_ids = np.where(np.arange(n) % 4 == 0, -1, np.arange(n) + 300)
            f.reference_frame_id, f.matched_scene_name = 11, 'scene'
        return f
    curr, last = frame(n0, False), frame(n1, True)
    config = {'localization': {'threshold': 12}}
    rt = ref_loc.tracker.Tracker.__new__(ref_loc.tracker.Tracker)
    rt.matcher, rt.config = CodeMatcher(), config
    for fail in (0, 10 ** 6):
        log_r, log_o = [], []
        ref_loc.pycolmap.absolute_pose_estimation = make_pose_fn(log_r, fail)
        ot = Tracker(None, CodeMatcher(), config, device='cpu', pose_fn=make_pose_fn(log_o, fail))
        a, b = rt.track_last_frame(curr, last), ot.track_last_frame(curr, last)
        _same(a, b)
        assert np.array_equal(log_r[0][0], log_o[0][0]) and log_r[0][2] == log_o[0][2]
    assert rt.matcher.calls[-1] == ot.matcher.calls[-1]
    # the bounding-box variant keeps the same bookkeeping (ids refer to the unfiltered current frame)
    ot = Tracker(None, CodeMatcher(), config, device='cpu', pose_fn=make_pose_fn([]))
    f = ot.track_last_frame_fast(curr, last)
    assert f['success'] and np.allclose(f['matched_keypoints'], curr.keypoints[f['matched_keypoint_ids'], :2])
    assert (f['matched_point3D_ids'] >= 0).all()
