"""Pins the CPU oracle: (1) against golden vectors generated from the reference's own modules
(oracle/make_golden.py), (2) when /root/reference is mounted, against those modules live."""
import numpy as np
import pytest
import torch

from oracle import pram_oracle as O, ref_loader as RL

needs_sfd2 = pytest.mark.skipif(RL.weight_path(RL.SFD2_WEIGHT) is None, reason='SFD2 checkpoint not staged')
needs_gml = pytest.mark.skipif(RL.weight_path(RL.GML_WEIGHT) is None, reason='GML checkpoint not staged')
needs_ref = pytest.mark.skipif(not RL.reference_available(), reason='reference tree not mounted')


def test_nms_and_selection_vs_golden(golden):
    g = golden('sfd2_160x120.npz')
    score = torch.from_numpy(g['score_map'])
    assert np.array_equal(O.simple_nms(score, 4).numpy(), g['nms4'])
    assert np.array_equal(O.simple_nms(score, 3).numpy(), g['nms3'])
    k, s = O.select_keypoints(torch.from_numpy(g['nms4'])[0], 0.005, int(g['min_keypoints']), int(g['max_keypoints']), 4)
    assert np.array_equal(k.numpy(), g['keypoints']) and np.array_equal(s.numpy(), g['scores'])
    k, s = O.select_keypoints(torch.from_numpy(g['nms4'])[0], 0.005, int(g['min_keypoints']), 4096, 4)
    assert np.array_equal(k.numpy(), g['keypoints_all']) and np.array_equal(s.numpy(), g['scores_all'])
    # row-major order in the take-all regime, score-descending in the top-k regime
    ka = g['keypoints_all']
    lin = ka[:, 1] * 160 + ka[:, 0]
    assert np.all(np.diff(lin) > 0)
    assert np.all(np.diff(g['scores']) <= 0)


@needs_sfd2
def test_sfd2_vs_golden(golden):
    g = golden('sfd2_160x120.npz')
    sd = RL.load_sfd2_state()
    torch.set_num_threads(8)
    out = O.sfd2_extract_local_global(sd, torch.from_numpy(g['image']),
                                      {'min_keypoints': int(g['min_keypoints']), 'max_keypoints': int(g['max_keypoints'])})
    assert np.allclose(out['score_map'].numpy(), g['score_map'], atol=1e-6)
    assert np.array_equal(out['keypoints'][0].numpy(), g['keypoints'])
    assert np.allclose(out['descriptors'][0].numpy(), g['descriptors'], atol=1e-5)
    sc, seg = O.sfd2_sample(out['score_map'], out['mid_features'], out['keypoints'][0], norm_desc=False)
    assert np.allclose(seg.numpy(), g['seg_descriptors'], atol=1e-4)
    assert np.allclose(sc.numpy(), g['sample_scores'], atol=1e-6)


def test_segnetvit_vs_golden(golden):
    g = golden('segnetvit_seed0.npz')
    sd = RL.random_segnetvit_state(int(g['n_class']), seed=int(g['seed']))
    pred = O.segnetvit_forward(sd, torch.from_numpy(g['seg_descriptors'])[None], torch.from_numpy(g['keypoints'])[None],
                               tuple(int(v) for v in g['image_shape']))
    assert np.allclose(pred[0].numpy(), g['prediction'], atol=2e-4)
    assert np.array_equal(pred[0].argmax(-1).numpy(), g['prediction'].argmax(-1))


def test_sinkhorn_and_matches_vs_golden(golden):
    g = golden('sinkhorn_70x93.npz')
    P = O.sinkhorn_with_dustbin(torch.from_numpy(g['dist']), torch.tensor(float(g['bin_score'])), 20)
    assert np.allclose(P.numpy(), g['P'], rtol=1e-5, atol=1e-7)
    i0, i1, s0, s1 = O.compute_matches(torch.from_numpy(g['P']), 0.2)
    assert np.array_equal(i0.numpy(), g['matches0']) and np.array_equal(i1.numpy(), g['matches1'])
    assert np.array_equal(s0.numpy(), g['scores0']) and np.array_equal(s1.numpy(), g['scores1'])
    assert (g['matches0'] > -1).sum() >= 30  # the planted correspondences are found


@needs_gml
def test_gml_vs_golden(golden):
    g = golden('gml_selfmatch.npz')
    sd = RL.load_gml_state()
    d0 = torch.from_numpy(g['descriptors0'])[None]
    k = torch.from_numpy(g['keypoints0'])
    perm = torch.from_numpy(g['perm'])
    data = {'descriptors0': d0, 'descriptors1': d0[:, perm], 'keypoints0': k[None], 'keypoints1': k[perm][None],
            'image_shape0': (1, 3, 160, 120), 'image_shape1': (1, 3, 160, 120)}
    out = O.gml_forward(sd, data)
    assert np.array_equal(out['matches0'][0].numpy(), g['matches0'])
    assert np.array_equal(out['matches1'][0].numpy(), g['matches1'])
    assert np.allclose(out['matching_scores0'][0].numpy(), g['scores0'], atol=1e-4)
    # known answer: keypoint perm[j] of set 0 is keypoint j of set 1
    m1 = g['matches1']
    ok = m1 > -1
    assert ok.sum() >= 0.9 * len(m1) and np.array_equal(m1[ok], g['perm'][ok])


@needs_ref
@needs_sfd2
def test_oracle_bit_identical_to_reference_modules():
    ref = RL.import_reference()
    torch.set_num_threads(8)
    sd = RL.load_sfd2_state()
    net = ref.sfd2.ResNet4x()
    net.load_state_dict(sd, strict=True)
    net.eval()
    img = O.frame_tensor(96, 128, seed=5)
    cfg = {'min_keypoints': 16, 'max_keypoints': 128}
    with torch.no_grad():
        r = net.extract_local_global({'image': img}, cfg)
    o = O.sfd2_extract_local_global(sd, img, cfg)
    assert torch.equal(r['score_map'], o['score_map'])
    assert torch.equal(r['keypoints'][0], o['keypoints'][0]) and torch.equal(r['scores'][0], o['scores'][0])
    assert torch.equal(r['descriptors'][0], o['descriptors'][0])
    asd = RL.random_gml_state(seed=3)
    g = ref.gml.GML({})
    g.load_state_dict(asd, strict=True)
    g.eval()
    k = r['keypoints'][0]
    d0 = r['descriptors'][0].t()[None]
    data = {'descriptors0': d0, 'descriptors1': d0.flip(1), 'keypoints0': k[None], 'keypoints1': k.flip(0)[None],
            'image0': img, 'image1': img}
    with torch.no_grad():
        rg = g(data)
    og = O.gml_forward(asd, data)
    assert torch.equal(rg['matches0'], og['matches0']) and torch.equal(rg['matching_scores0'], og['matching_scores0'])


def test_pose_oracle_known_answer():
    """absolute_pose_estimation is parity-unpinned (pycolmap absent); pin it on synthetic known poses."""
    rs = np.random.RandomState(0)
    n = 200
    ang = 0.2
    R = O.quat_to_rotmat(np.array([np.cos(ang / 2), 0, np.sin(ang / 2), 0.0]))
    t = np.array([0.1, -0.2, 0.3])
    Xc = np.stack([rs.uniform(-1, 1, n), rs.uniform(-1, 1, n), rs.uniform(2, 5, n)], 1)
    X = (Xc - t) @ R  # X_cam = R X + t
    f, cx, cy = 525.0, 320.0, 240.0
    uv = np.stack([f * Xc[:, 0] / Xc[:, 2] + cx, f * Xc[:, 1] / Xc[:, 2] + cy], 1) + rs.normal(0, 0.5, (n, 2))
    out_idx = rs.choice(n, 40, replace=False)
    uv[out_idx] += rs.uniform(50, 100, (40, 2))
    cam = {'model': 'SIMPLE_PINHOLE', 'width': 640, 'height': 480, 'params': [f, cx, cy]}
    ret = O.absolute_pose_estimation(uv, X, cam, max_error=8.0, max_num_trials=2000)
    assert ret is not None
    q_gt = O.rotmat_to_quat(R)
    e_r, e_t = O.pose_error(ret['qvec'], ret['tvec'], q_gt, t)
    assert e_r < 0.5 and e_t < 0.05
    gt_inl = np.ones(n, bool)
    gt_inl[out_idx] = False
    assert (ret['inliers'] == gt_inl).mean() > 0.98


@needs_ref
def test_next_row_oracles_vs_reference_methods():
    """add_segmentations / process_segmentations restatements vs the reference's own methods, imported with
    stub modules for the dependencies that are absent here (pycolmap, h5py, ...)."""
    import sys
    import types
    RL.import_reference()
    class _Any(types.ModuleType):  # stub module: any attribute resolves to a dummy class
        def __getattr__(self, item):
            if item.startswith('__'):
                raise AttributeError(item)
            return type(item, (), {})
    for name in ('pycolmap', 'h5py', 'progressbar', 'open3d', 'tensorboardX', 'pypangolin', 'OpenGL', 'OpenGL.GL'):
        sys.modules.setdefault(name, _Any(name))
    try:
        from localization.frame import Frame
        from localization.multimap3d import MultiMap3D
    except Exception as e:  # noqa: BLE001 -- the reference pulls many optional dependencies
        pytest.skip(f'reference localization modules not importable here: {e!r}')
    g = torch.Generator().manual_seed(1)
    logits = torch.randn(500, 113, generator=g) * 3
    logits[:200, 0] += 9
    fr = Frame.__new__(Frame)
    fr.keypoints = np.zeros((500, 3), np.float32)
    fr.descriptors = np.zeros((500, 128), np.float32)
    fr.initialize_localization_variables = lambda: None
    fr.add_segmentations(logits, 0.95)
    keep, scores, ids = O.add_segmentations(logits, 0.95)
    assert fr.keypoints.shape[0] == int(keep.sum())
    assert np.array_equal(fr.seg_ids, ids.numpy()) and np.allclose(fr.seg_scores, scores.numpy())
    mm = MultiMap3D.__new__(MultiMap3D)
    ref = mm.process_segmentations(torch.from_numpy(fr.segmentations), topk=20)
    ours = O.process_segmentations(torch.from_numpy(fr.segmentations), topk=20)
    assert len(ref) == len(ours)
    for (s0, i0, v0), (s1, i1, v1) in zip(ref, ours):
        assert s0 == s1 and np.array_equal(i0, i1) and v0 == v1


@needs_ref
def test_nearest_neighbor_oracle_vs_reference_module():
    """The NN-matcher restatement against the reference's own plugin (importable as-is)."""
    RL.import_reference()
    from localization.matchers.nearest_neighbor import NearestNeighbor
    g = torch.Generator().manual_seed(5)
    d0 = torch.nn.functional.normalize(torch.randn(2, 128, 300, generator=g), dim=1)
    d1 = torch.nn.functional.normalize(torch.randn(2, 128, 257, generator=g), dim=1)
    d1[:, :, :100] = torch.nn.functional.normalize(d0[:, :, 50:150] + 0.1 * torch.randn(2, 128, 100, generator=g), dim=1)
    for conf in ({}, {'ratio_threshold': 0.9}, {'distance_threshold': 0.7, 'do_mutual_check': False},
                 {'ratio_threshold': 0.95, 'distance_threshold': 0.9}):
        ref = NearestNeighbor(conf)({'descriptors0': d0, 'descriptors1': d1})
        ours = O.nearest_neighbor_forward(d0, d1, **{**NearestNeighbor.default_conf, **conf})
        assert torch.equal(ref['matches0'], ours['matches0'])
        assert torch.equal(ref['matching_scores0'], ours['matching_scores0'])


@needs_ref
def test_adagml_oracle_vs_reference_module():
    """O.adagml_forward against the reference's own ``nets.adagml.AdaGML.produce_matches`` (CPU shim of SURVEY.md 8c:
    nets.adagml.sink_algorithm = nets.gml.sink_algorithm) on a case that really prunes tokens over several layers and
    exits early: identical pruning trace, identical matches, scores equal to the last float32 bit (torch's threaded CPU
    reductions are allowed one ulp of run-to-run noise)."""
    ref = RL.import_reference()
    sd = RL.calibrated_adagml_state()
    net = ref.adagml.AdaGML({})
    net.load_state_dict(sd, strict=True)
    net.eval()
    g = torch.Generator().manual_seed(0)
    m = n = 400
    d0 = torch.nn.functional.normalize(torch.randn(1, m, 128, generator=g), dim=-1)
    perm = torch.randperm(n, generator=g)
    d1 = d0[:, perm] + 0.02 * torch.randn(1, n, 128, generator=g)
    k0 = torch.rand(1, m, 2, generator=g) * torch.tensor([640., 480.])
    data = {'descriptors0': d0, 'descriptors1': d1, 'keypoints0': k0, 'keypoints1': k0[:, perm],
            'scores0': torch.rand(1, m, generator=g), 'scores1': torch.rand(1, n, generator=g),
            'image_shape0': (1, 3, 640, 480), 'image_shape1': (1, 3, 640, 480)}
    with torch.no_grad():
        r = net.produce_matches(data)
    o = O.adagml_forward(sd, data, return_trace=True)
    assert len(o['trace']) >= 3 and o['trace'][-1][1] < 200 and o['last_layer'] < 8   # pruned and stopped early
    assert torch.equal(r['matches0'], o['matches0'])
    assert (r['matching_scores0'] > 0).sum() > 50   # the surviving tokens carry real Sinkhorn scores (seeded weights: no match above 0.2)
    assert torch.allclose(r['matching_scores0'], o['matching_scores0'], rtol=0, atol=1e-6)
