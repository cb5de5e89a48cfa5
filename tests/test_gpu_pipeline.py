"""Parity of the BENCHED path: ``LocalizationPipeline.localize`` at the BASELINE.json configuration
(640x480, K = 1024, shipped SFD2 + GML checkpoints, seeded SegNetViT with 113 classes, PnP max_error 8)
against the CPU oracle run frame by frame, plus the execution-mode invariants the bench relies on:
CUDA-graph replay == eager, two streams == one stream, batch 32 == batch 4 on the shared frames.

Protocol per frame:
* keypoints: the GPU set against the oracle's own selection under the margin protocol (every disagreement must
  be explained by a score margin below 4x the score-map tolerance);
* every later stage: the oracle is run on the GPU's keypoint set (its own maps, its own sampling), so the
  token sets are identical and logits / labels / matching scores / matches / pose compare one to one.
"""
import json
import os
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import pram_oracle as O, ref_loader as RL

pytestmark = pytest.mark.gpu

H, W, K, NCLASS = 480, 640, 1024, 113
FOCAL, MAX_ERROR = 525.0, 8.0
SCORE_TOL = 2e-4     # bf16x3 score map (tests/test_gpu_nets.py)
LOGIT_TOL = 1e-2     # 15 layers at 1024 tokens, bf16x3, inputs sampled from maps that differ by <= 5e-4 of their max
MSCORE_TOL = 5e-3    # matching scores (as in test_gml_vs_golden)
REPORT = Path(__file__).resolve().parents[1] / 'gpurun_out'


def _build(dev, precision='bf16x3'):
    from pram_b200.nets.sfd2 import ResNet4x
    from pram_b200.nets.segnetvit import SegNetViT
    from pram_b200.nets.gml import GML
    from pram_b200.runner import LocalizationPipeline
    sd_sfd2, sd_gml = RL.load_sfd2_state(), RL.load_gml_state()
    if sd_sfd2 is None or sd_gml is None:
        pytest.skip('shipped checkpoints not staged')
    sd_vit = RL.random_segnetvit_state(NCLASS, seed=0)
    sfd2 = ResNet4x(); sfd2.load_state_dict(sd_sfd2, strict=True)
    vit = SegNetViT({'n_class': NCLASS, 'n_layers': 15, 'output_dim': 1024, 'descriptor_dim': 256})
    vit.load_state_dict(sd_vit, strict=True)
    gml = GML({}); gml.load_state_dict(sd_gml, strict=True)
    if precision == 'bf16x3+f16desc':   # the mixed mode: single-pass fp16 descriptor head, everything else bf16x3
        sfd2.set_precision('bf16x3', 'f16'); vit.set_precision('bf16x3', 'f16'); gml.set_precision('bf16x3', 'f16')
    else:
        for m in (sfd2, vit, gml):
            m.set_precision(precision)
    pipe = LocalizationPipeline(sfd2, vit, gml, max_keypoints=K, focal=FOCAL, ransac_max_error=MAX_ERROR, device=dev)
    return pipe, (sd_sfd2, sd_vit, sd_gml)


def _frames(n, seed0=0):
    return torch.cat([O.frame_tensor(H, W, seed=seed0 + i) for i in range(n)], 0)


@pytest.fixture(scope='module')
def case(lib, dev):
    pipe, sds = _build(dev)
    frames = _frames(4)
    fd = frames.to(dev)
    with torch.no_grad():
        smap = pipe.build_synthetic_map(fd, seed=0)
        out = pipe.localize(fd, smap)
    torch.cuda.synchronize()
    return {'pipe': pipe, 'sds': sds, 'frames': frames, 'fd': fd, 'smap': smap, 'out': out}


def _check_frame_against_oracle(i, frames, out, smap, sds, report, dims=None):
    """Frame ``i`` of a pipeline result against the oracle; appends the measured deviations to ``report``."""
    H, W, K, FOCAL = dims or (globals()['H'], globals()['W'], globals()['K'], globals()['FOCAL'])
    sd_sfd2, sd_vit, sd_gml = sds
    img = frames[i:i + 1]
    torch.set_num_threads(os.cpu_count() or 8)
    with torch.no_grad():
        ref = O.sfd2_extract_local_global(sd_sfd2, img, {'min_keypoints': 128, 'max_keypoints': K})
    n = int(out['num_keypoints'][i])
    kp = out['keypoints'][i, :n].cpu()
    # ---- keypoints: margin protocol against the oracle's own selection ----
    ours = {(float(x), float(y)) for x, y in kp}
    theirs = {(float(x), float(y)) for x, y in ref['keypoints'][0]}
    assert n == len(ours) and len(ours & theirs) >= 0.97 * len(theirs), (n, len(ours & theirs), len(theirs))
    s = ref['score_map'][0].numpy()
    kth = float(ref['scores'][0].min()) if len(theirs) == K else 0.005
    for x, y in ours ^ theirs:
        yy, xx = int(y), int(x)
        win = s[max(0, yy - 4):yy + 5, max(0, xx - 4):xx + 5]
        second = np.sort(win.ravel())[-2]
        margin = min(abs(s[yy, xx] - 0.005), abs(s[yy, xx] - second), abs(s[yy, xx] - kth))
        assert margin < 4 * SCORE_TOL, (x, y, margin)
    # ---- recognition on the identical token set ----
    with torch.no_grad():
        _, seg = O.sfd2_sample(ref['score_map'], ref['mid_features'], kp, norm_desc=False)
        logits = O.segnetvit_forward(sd_vit, seg.t()[None], kp[None], img.shape)[0]
    pred = out['prediction'][i, :n].cpu()
    dl = (pred - logits).abs().max().item()
    assert dl < LOGIT_TOL, dl
    top2 = torch.sort(logits, -1).values[:, -2:]
    decisive = (top2[:, 1] - top2[:, 0]) > 2 * LOGIT_TOL
    assert decisive.float().mean() > 0.5
    assert torch.equal(out['labels'][i, :n].cpu()[decisive], logits.argmax(-1)[decisive])
    # ---- matching against the synthetic map, identical inputs except the query descriptors (oracle-sampled) ----
    with torch.no_grad():
        d0 = O.sample_map(kp, ref['desc_map'], 4, True).t()[None]
        dd = (out_desc(out, i, n) - d0[0]).abs().max().item() if 'descriptors' in out else None
        nr = int(smap.num[i]) if smap.num is not None else smap.descriptors.shape[1]   # real reference keypoints (rest: padding)
        mr = O.gml_forward(sd_gml, {'descriptors0': d0, 'keypoints0': kp[None],
                                    'descriptors1': smap.descriptors[i:i + 1, :nr].cpu(), 'keypoints1': smap.keypoints[i:i + 1, :nr].cpu(),
                                    'image_shape0': (1, 3, W, H), 'image_shape1': (1, 3, W, H)})
    s0 = out['matching_scores0'][i, :n].cpu()
    ds = (s0 - mr['matching_scores0'][0]).abs().max().item()
    assert ds < MSCORE_TOL, ds
    dec = (mr['matching_scores0'][0] - 0.2).abs() > 2 * MSCORE_TOL
    m0 = out['matches0'][i, :n].cpu()
    assert torch.equal(m0[dec], mr['matches0'][0][dec])
    # known answer of the synthetic map: reference j <- query perm[j].  (Outlier references keep their 2-D position,
    # so the matcher still pairs many of them; their 3-D points are wrong and PnP must reject them -- checked below.)
    perm, outl = smap.perm[i].cpu(), smap.outlier[i].cpu()
    ok = m0 > -1
    correct = perm[m0[ok]] == torch.nonzero(ok)[:, 0]
    assert ok.sum() > 0.6 * n and correct.float().mean() > 0.97
    # ---- pose: against the oracle's estimate from ITS matches, against the known pose, and self-consistency ----
    mo = mr['matches0'][0].numpy()
    sel = mo > -1
    cam = {'model': 'PINHOLE', 'width': W, 'height': H, 'params': [FOCAL, FOCAL, W / 2.0, H / 2.0]}
    xyz = smap.xyz[i].double().cpu().numpy()
    ret = O.absolute_pose_estimation(kp.numpy()[sel].astype(np.float64) + 0.5, xyz[mo[sel]], cam, max_error=MAX_ERROR,
                                     max_num_trials=2000)
    assert ret is not None and bool(out['pose_success'][i])
    q, t = out['qvec'][i].cpu().numpy(), out['tvec'][i].cpu().numpy()
    e_r, e_t = O.pose_error(q, t, ret['qvec'], ret['tvec'])
    assert e_r < 0.3 and e_t < 0.03, (e_r, e_t)
    g_r, g_t = O.pose_error(q, t, O.rotmat_to_quat(smap.R[i].double().cpu().numpy()), smap.t[i].double().cpu().numpy())
    assert g_r < 0.3 and g_t < 0.03, (g_r, g_t)
    inl = out['inliers'][i, :n].cpu().numpy()
    x = (kp.numpy().astype(np.float64) + 0.5 - np.array([W / 2.0, H / 2.0])) / FOCAL
    mm = m0.numpy()
    e = O._reproj_sq_err(O.quat_to_rotmat(q), t, x[mm > -1], xyz[mm[mm > -1]])
    exp = np.zeros(n, bool)
    exp[mm > -1] = e <= (MAX_ERROR / FOCAL) ** 2
    assert (inl != exp).sum() <= 2, (inl != exp).sum()      # a residual within 1 ulp of the threshold may flip
    assert int(out['num_inliers'][i]) == int(inl.sum())
    assert not inl[mm > -1][outl[mm[mm > -1]].numpy()].any()  # no planted outlier survives as an inlier
    assert abs(int(out['num_inliers'][i]) - ret['num_inliers']) <= 0.02 * n
    report.append({'frame': i, 'n': n, 'kpt_common': len(ours & theirs), 'logit_maxdiff': dl, 'desc_maxdiff': dd,
                   'mscore_maxdiff': ds, 'matched': int(ok.sum()), 'rot_err_vs_oracle_deg': e_r, 't_err_vs_oracle_m': e_t,
                   'rot_err_vs_known_deg': g_r, 't_err_vs_known_m': g_t, 'inliers': int(inl.sum()),
                   'oracle_inliers': int(ret['num_inliers'])})


def out_desc(out, i, n):
    return out['descriptors'][i, :n].cpu()


def _dump(name, report):
    try:
        REPORT.mkdir(exist_ok=True)
        (REPORT / name).write_text(json.dumps(report, indent=1))
    except OSError:
        pass


def test_pipeline_batch4_vs_oracle(case):
    report = []
    for i in range(4):
        _check_frame_against_oracle(i, case['frames'], case['out'], case['smap'], case['sds'], report)
    _dump('pipeline_parity_b4.json', report)


def test_pipeline_mixed_precision_vs_oracle(case, dev):
    """The mixed mode (descriptor head in single-pass fp16, attention probabilities as one fp16 plane, everything that
    feeds keypoint selection in bf16x3) must pass the SAME oracle comparison with the SAME tolerances; its keypoints are
    bit-identical to the parity mode's (the detector branch is untouched)."""
    pipe, _ = _build(dev, 'bf16x3+f16desc')
    with torch.no_grad():
        out = pipe.localize(case['fd'], case['smap'])
    torch.cuda.synchronize()
    assert torch.equal(out['keypoints'], case['out']['keypoints'])
    assert not torch.equal(out['descriptors'], case['out']['descriptors'])   # the fp16 head really ran
    assert not torch.equal(out['prediction'], case['out']['prediction'])     # ... and the fp16 attention probabilities
    assert (out['prediction'] - case['out']['prediction']).abs().max() < 2e-3
    report = []
    for i in range(4):
        _check_frame_against_oracle(i, case['frames'], out, case['smap'], case['sds'], report)
    _dump('pipeline_parity_b4_mixed.json', report)


EXACT_KEYS = ('keypoints', 'num_keypoints', 'prediction', 'labels', 'seg_ids', 'non_bg', 'matches0', 'matches1',
              'matching_scores0', 'matching_scores1', 'qvec', 'tvec', 'num_inliers', 'inliers', 'pose_success')


def _assert_same(a, b, keys=EXACT_KEYS):
    for k in keys:
        assert torch.equal(a[k], b[k]), k
    for k in a['landmarks']:
        assert torch.equal(a['landmarks'][k], b['landmarks'][k]), ('landmarks', k)


def test_graph_replay_equals_eager(case):
    """capture()/replay() is what bench.py times: it must reproduce the eager result bit for bit, and keep doing so
    when the static input buffer is refilled with other frames."""
    pipe, fd, smap = case['pipe'], case['fd'], case['smap']
    eager = {k: (v.clone() if torch.is_tensor(v) else {kk: vv.clone() for kk, vv in v.items()}) for k, v in case['out'].items()}
    n = pipe.capture(fd, smap)
    assert n > 50
    rep = pipe.replay()
    torch.cuda.synchronize()
    _assert_same(eager, rep)
    # refill the static input with the frames in another order and back: results follow the input
    flipped = pipe.replay(fd.flip(0))
    torch.cuda.synchronize()
    assert torch.equal(flipped['keypoints'], eager['keypoints'].flip(0))
    assert torch.equal(flipped['prediction'], eager['prediction'].flip(0))
    rep = pipe.replay(fd)
    torch.cuda.synchronize()
    _assert_same(eager, rep)


def test_two_streams_equals_one_stream(case, dev):
    pipe, fd, smap = case['pipe'], case['fd'], case['smap']
    saved = pipe.two_streams
    try:
        pipe.two_streams = False
        with torch.no_grad():
            one = pipe.localize(fd, smap)
        torch.cuda.synchronize()
        pipe.two_streams = True
        with torch.no_grad():
            two = pipe.localize(fd, smap)
        torch.cuda.synchronize()
    finally:
        pipe.two_streams = saved
    _assert_same(one, two)
    _assert_same(one, case['out'])


def test_programmatic_dependent_launch_modes_are_equivalent(case, lib):
    """pram_set_pdl: 0 = plain stream order, 1 = dependent kernels scheduled during the predecessor's drain (default),
    2 = released as soon as the predecessor's CTAs are resident.  Only the schedule may change: eager runs and a
    freshly captured graph must give bit-identical results in every mode."""
    pipe, fd, smap = case['pipe'], case['fd'], case['smap']
    saved = lib.pram_get_pdl()
    try:
        for mode in (0, 2, 1):
            assert lib.pram_set_pdl(mode) == 0 and lib.pram_get_pdl() == mode
            with torch.no_grad():
                out = pipe.localize(fd, smap)
            torch.cuda.synchronize()
            _assert_same(case['out'], out)
            pipe.capture(fd, smap)
            for _ in range(3):
                rep = pipe.replay()
            torch.cuda.synchronize()
            _assert_same(case['out'], rep)
    finally:
        lib.pram_set_pdl(saved)


def test_pipeline_batch32_vs_oracle_and_batch4(case, dev):
    """The bench configuration itself (32 frames per step): frames 0-3 must equal the batch-4 run bit for bit (every
    output element is produced by the same instruction sequence whatever the batch), and three frames spread over the
    batch are checked against the oracle."""
    pipe, sds = case['pipe'], case['sds']
    frames = _frames(32)
    fd = frames.to(dev)
    with torch.no_grad():
        smap = pipe.build_synthetic_map(fd, seed=0)
        out = pipe.localize(fd, smap)
    torch.cuda.synchronize()
    assert (out['num_keypoints'] == K).all()  # the bench frames all fill the keypoint budget (no padded tokens)
    for k in ('keypoints', 'num_keypoints', 'prediction', 'labels'):
        assert torch.equal(out[k][:4], case['out'][k]), k
    report = []
    for i in (5, 17, 31):
        _check_frame_against_oracle(i, frames, out, smap, sds, report)
    _dump('pipeline_parity_b32.json', report)


def test_padded_batch_equals_per_frame_reference(lib, dev):
    """Frames with FEWER keypoints than the budget, batched in the fixed [B, K] layout (advisor finding of round 1: padded
    slots took part in attention / Sinkhorn / PnP).  160x120 frames yield ~200-300 keypoints against K = 512, and a
    different count per frame; with the counts threaded through attention (key masking), Sinkhorn (per-pair problem
    size) and PnP, every frame must equal the oracle run on that frame alone with exactly its n[b] keypoints -- same
    checker and tolerances as the full-budget test."""
    Hs, Ws, Ks, Fs = 120, 160, 512, 150.0
    from pram_b200.runner import LocalizationPipeline
    pipe, sds = _build(dev)
    pipe = LocalizationPipeline(pipe.sfd2, pipe.segnet, pipe.matcher, max_keypoints=Ks, focal=Fs, ransac_max_error=MAX_ERROR, device=dev)
    pipe.cfg = {'min_keypoints': 32, 'max_keypoints': Ks}
    frames = torch.cat([O.frame_tensor(Hs, Ws, seed=20 + i) for i in range(3)], 0)
    fd = frames.to(dev)
    with torch.no_grad():
        smap = pipe.build_synthetic_map(fd, seed=1)
        out = pipe.localize(fd, smap)
    torch.cuda.synchronize()
    n = out['num_keypoints'].cpu()
    assert (n < Ks).all() and len(set(n.tolist())) > 1 and torch.equal(smap.num.cpu(), n)
    for i in range(3):
        ni = int(n[i])
        assert (out['matches0'][i, ni:] == -1).all() and (out['matching_scores0'][i, ni:] == 0).all()
        assert not out['inliers'][i, ni:].any() and not out['non_bg'][i, ni:].any()
        assert (out['matches0'][i, :ni] < ni).all()          # nothing is matched to a padded reference slot
    report = []
    for i in range(3):
        _check_frame_against_oracle_small(i, frames, out, smap, sds, report, (Hs, Ws, Ks, Fs))
    _dump('pipeline_parity_padded.json', report)


def _check_frame_against_oracle_small(i, frames, out, smap, sds, report, dims):
    """Same checker; the oracle's selection uses min_keypoints = 32 like the pipeline of this test."""
    orig = O.sfd2_extract_local_global
    try:
        O.sfd2_extract_local_global = lambda sd, img, cfg: orig(sd, img, {**cfg, 'min_keypoints': 32})
        _check_frame_against_oracle(i, frames, out, smap, sds, report, dims)
    finally:
        O.sfd2_extract_local_global = orig
