"""GPU parity of the network stages against the CPU oracle and the reference-generated golden vectors."""
import numpy as np
import pytest
import torch

from oracle import pram_oracle as O, ref_loader as RL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops(lib, dev):
    from pram_b200 import ops as _ops
    return _ops


PRECISIONS = ['fp32', 'bf16x3']


def _sfd2(dev, sd, precision='bf16x3'):
    from pram_b200.nets.sfd2 import ResNet4x
    net = ResNet4x()
    net.load_state_dict(sd, strict=True)
    return net.to(dev).set_precision(precision)


@pytest.mark.parametrize('precision', PRECISIONS + ['bf16'])
@pytest.mark.parametrize('hw', [(72, 88), (61, 83)])
def test_conv_stack_random_weights(lib, dev, precision, hw):
    """Conv stack vs torch-CPU fp32, relative to each map's max.  fp32 CUDA cores: 2e-4 (accumulation
    order over up to 2304 products); bf16x3 tensor cores: 5e-4 (16-bit effective mantissa, 21 layers);
    plain bf16: 8e-2 (reported, not a parity mode).  The odd size exercises ragged tiles and the
    phase-split layout with odd H/W."""
    sd = RL.random_sfd2_state(seed=1)
    img = torch.randn(2, 3, *hw, generator=torch.Generator().manual_seed(0))
    ref = O.sfd2_trunk(sd, img)
    net = _sfd2(dev, sd, precision)
    t = net._trunk(img.to(dev))
    tol = {'fp32': 2e-4, 'bf16x3': 5e-4, 'bf16': 8e-2}[precision]
    for name, key in (('out1b', 'out1b'), ('out2b', 'out2b'), ('out3b', 'out3b'), ('out4', 'out4'),
                      ('logits', 'logits'), ('desc', 'desc_map')):
        a = net._nchw(t[name]).cpu()
        b = ref[key]
        assert a.shape == b.shape, name
        err = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-6)
        assert err < tol, (name, err)


def test_grouped_conv_odd_width(ops, dev):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 256, 9, 13, generator=g)
    w = torch.randn(256, 8, 3, 3, generator=g) * 0.1
    b = torch.randn(256, generator=g)
    ref = torch.relu(torch.nn.functional.conv2d(x, w, b, padding=1, groups=32))
    wp = w.view(32, 8, 8, 3, 3).permute(3, 4, 2, 1, 0).reshape(9, 8, 8, 32).contiguous()
    out = ops.gconv3x3_f32(x.permute(0, 2, 3, 1).contiguous().to(dev), wp.to(dev), b.to(dev), True)
    assert torch.allclose(out.permute(0, 3, 1, 2).cpu(), ref, rtol=1e-4, atol=1e-5)


@pytest.mark.skipif(RL.weight_path(RL.SFD2_WEIGHT) is None, reason='SFD2 checkpoint not staged')
@pytest.mark.parametrize('precision', PRECISIONS)
def test_extract_local_global_shipped_weights(lib, dev, golden, precision):
    """End to end with the shipped SFD2 checkpoint on the golden frame.  Score map within 2e-5; keypoint
    indices exact under the margin protocol: a reference keypoint may be missing only if its score is
    within the score-map tolerance of the threshold / k-th score / a 9x9 neighbour."""
    g = golden('sfd2_160x120.npz')
    net = _sfd2(dev, RL.load_sfd2_state(), precision)
    img = torch.from_numpy(g['image']).to(dev)
    out = net.extract_local_global({'image': img}, {'min_keypoints': 32, 'max_keypoints': 4096})
    sm = out['score_map'].cpu().numpy()
    tol = {'fp32': 2e-5, 'bf16x3': 2e-4}[precision]
    assert np.abs(sm - g['score_map']).max() < tol, np.abs(sm - g['score_map']).max()
    ours = {(float(x), float(y)) for x, y in out['keypoints'][0].cpu()}
    theirs = {(float(x), float(y)) for x, y in g['keypoints_all']}
    common = ours & theirs
    assert len(common) >= 0.97 * len(theirs)
    for x, y in ours ^ theirs:  # every disagreement must be explained by a margin below tolerance
        s = g['score_map'][0]
        yy, xx = int(y), int(x)
        win = s[max(0, yy - 4):yy + 5, max(0, xx - 4):xx + 5]
        second = np.sort(win.ravel())[-2]
        margin = min(abs(s[yy, xx] - 0.005), abs(s[yy, xx] - second))
        assert margin < 4 * tol, (x, y, margin)
    # descriptors of the common keypoints
    idx_o = {(float(x), float(y)): i for i, (x, y) in enumerate(out['keypoints'][0].cpu())}
    idx_r = {(float(x), float(y)): i for i, (x, y) in enumerate(g['keypoints_all'])}
    d_o = out['descriptors'][0].cpu().numpy()
    for kxy in list(common)[:50]:
        assert np.abs(d_o[:, idx_o[kxy]] - g['descriptors_all'][:, idx_r[kxy]]).max() < 10 * tol
    # API shapes of the reference contract
    assert out['desc_map'].shape == (1, 128, 30, 40) and out['mid_features'].shape == (1, 256, 30, 40)
    assert len(out['global_descriptors']) == 4 and out['descriptors'][0].shape[0] == 128
    sc, seg = net.sample(out['score_map'], out['mid_features'], torch.from_numpy(g['keypoints']).to(dev), norm_desc=False)
    assert np.abs(seg.cpu().numpy() - g['seg_descriptors']).max() < 100 * tol
    assert np.abs(sc.cpu().numpy() - g['sample_scores']).max() < tol


def test_attention_vs_torch(ops, dev):
    g = torch.Generator().manual_seed(0)
    b, h, nq, nk = 2, 4, 75, 130
    q, k, v = (torch.randn(b, h, n, 64, generator=g) for n in (nq, nk, nk))
    attn = torch.softmax(torch.einsum('bhid,bhjd->bhij', q, k) * 0.125, -1)
    ref = torch.einsum('bhij,bhjd->bhid', attn, v).transpose(1, 2).flatten(-2)
    out = torch.empty(b, nq, 256, device=dev)
    cm = torch.empty(b, nk, device=dev)
    ops.attention_f32(q.to(dev).contiguous(), k.to(dev).contiguous(), v.to(dev).contiguous(), b, h, nq, nk, 0.125, out, 256, cm)
    assert torch.allclose(out.cpu(), ref, rtol=1e-4, atol=1e-5)
    assert torch.allclose(cm.cpu(), attn.mean(1).mean(1), rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize('precision', PRECISIONS)
def test_segnetvit_vs_golden(lib, dev, golden, precision):
    from pram_b200.nets.segnetvit import SegNetViT
    g = golden('segnetvit_seed0.npz')
    sd = RL.random_segnetvit_state(int(g['n_class']), seed=int(g['seed']))
    m = SegNetViT({'n_class': int(g['n_class']), 'n_layers': 15, 'output_dim': 1024, 'descriptor_dim': 256})
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).set_precision(precision)
    shape = tuple(int(v) for v in g['image_shape'])
    x = torch.from_numpy(g['seg_descriptors'])[None].to(dev)
    k = torch.from_numpy(g['keypoints'])[None].to(dev)
    pred = m({'seg_descriptors': x, 'keypoints': k, 'image': torch.empty(shape, device='meta')})['prediction'][0].cpu().numpy()
    # 15 layers: fp32 re-ordered sums -> 1e-3 absolute on logits of magnitude ~1; bf16x3 -> 5e-3
    tol = {'fp32': 1e-3, 'bf16x3': 5e-3}[precision]
    assert np.abs(pred - g['prediction']).max() < tol, np.abs(pred - g['prediction']).max()
    top2 = np.sort(g['prediction'], -1)[:, -2:]
    decisive = (top2[:, 1] - top2[:, 0]) > 2 * tol
    assert np.array_equal(pred.argmax(-1)[decisive], g['prediction'].argmax(-1)[decisive])
    # batching is semantically safe (reference: B=4 vs B=1 identical arg-max)
    pred2 = m({'seg_descriptors': x.repeat(3, 1, 1), 'keypoints': k.repeat(3, 1, 1), 'image': torch.empty(shape, device='meta')})['prediction']
    assert torch.allclose(pred2[2].cpu(), torch.from_numpy(pred), atol=1e-5)


def test_sinkhorn_match_vs_golden(ops, dev, golden):
    g = golden('sinkhorn_70x93.npz')
    dist = torch.from_numpy(g['dist']).to(dev)
    for cluster in (1, 2, 8):
        m0, m1, s0, s1, P = ops.sinkhorn_match(dist, torch.tensor(float(g['bin_score']), device=dev), 20, 0.2,
                                               cluster=cluster, return_P=True)
        assert np.allclose(P.cpu().numpy(), g['P'], rtol=2e-4, atol=1e-7), cluster
        assert np.array_equal(m0.cpu().numpy(), g['matches0']) and np.array_equal(m1.cpu().numpy(), g['matches1'])
        assert np.allclose(s0.cpu().numpy(), g['scores0'], rtol=2e-4, atol=1e-7)
        assert np.allclose(s1.cpu().numpy(), g['scores1'], rtol=2e-4, atol=1e-7)


@pytest.mark.parametrize('m,n', [(1, 1), (5, 300), (1024, 1024), (300, 2048), (130, 4096)])
def test_sinkhorn_match_shapes_vs_oracle(ops, dev, m, n):
    g = torch.Generator().manual_seed(m * 7 + n)
    dist = torch.randn(1, m, n, generator=g) * 2
    for i in range(min(m, n) // 2):
        dist[0, i, (i * 3) % n] += 15
    bin_score = torch.tensor(0.7)
    P = O.sinkhorn_with_dustbin(dist, bin_score, 20)
    i0, i1, s0, s1 = O.compute_matches(P, 0.2)
    m0, m1, t0, t1, Pg = ops.sinkhorn_match(dist.to(dev), bin_score.to(dev), 20, 0.2, return_P=True)
    assert torch.allclose(Pg.cpu(), P, rtol=5e-4, atol=1e-7)
    # total mass agrees with the oracle (no size-independent marginal property exists: with M != N the
    # reference's marginals are inconsistent and the 20 iterations do not converge to either of them)
    assert abs(Pg.sum().item() - P.sum().item()) < 1e-3 * P.sum().item()
    decisive = (s0[0] - 0.2).abs() > 1e-3
    assert torch.equal(m0.cpu()[0][decisive], i0[0][decisive])
    assert torch.allclose(t0.cpu(), s0, rtol=5e-4, atol=1e-6)


@pytest.mark.skipif(RL.weight_path(RL.GML_WEIGHT) is None, reason='GML checkpoint not staged')
@pytest.mark.parametrize('precision', PRECISIONS)
def test_gml_vs_golden(lib, dev, golden, precision):
    from pram_b200.nets.gml import GML
    g = golden('gml_selfmatch.npz')
    net = GML({})
    net.load_state_dict(RL.load_gml_state(), strict=True)
    net = net.to(dev).set_precision(precision)
    d0 = torch.from_numpy(g['descriptors0'])[None].to(dev)
    k = torch.from_numpy(g['keypoints0']).to(dev)
    perm = torch.from_numpy(g['perm']).to(dev)
    data = {'descriptors0': d0, 'descriptors1': d0[:, perm], 'keypoints0': k[None], 'keypoints1': k[perm][None],
            'image_shape0': (1, 3, 160, 120), 'image_shape1': (1, 3, 160, 120)}
    out = net(data)
    s0 = out['matching_scores0'][0].cpu().numpy()
    assert np.abs(s0 - g['scores0']).max() < 5e-3
    decisive0 = np.abs(g['scores0'] - 0.2) > 1e-2
    assert np.array_equal(out['matches0'][0].cpu().numpy()[decisive0], g['matches0'][decisive0])
    decisive1 = np.abs(g['scores1'] - 0.2) > 1e-2
    assert np.array_equal(out['matches1'][0].cpu().numpy()[decisive1], g['matches1'][decisive1])
    assert out['matches0'].dtype == torch.int64


@pytest.mark.parametrize('precision', PRECISIONS)
def test_gml_random_weights_batched_vs_oracle(lib, dev, precision):
    from pram_b200.nets.gml import GML
    sd = RL.random_gml_state(seed=5)
    net = GML({})
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).set_precision(precision)
    g = torch.Generator().manual_seed(0)
    b, m, n = 2, 90, 70
    d0 = torch.nn.functional.normalize(torch.randn(b, m, 128, generator=g), dim=-1)
    d1 = torch.nn.functional.normalize(torch.randn(b, n, 128, generator=g), dim=-1)
    d1[:, :50] = d0[:, 20:70] + 0.01 * torch.randn(b, 50, 128, generator=g)
    k0 = torch.rand(b, m, 2, generator=g) * torch.tensor([640., 480.])
    k1 = torch.rand(b, n, 2, generator=g) * torch.tensor([640., 480.])
    k1[:, :50] = k0[:, 20:70]
    data = {'descriptors0': d0, 'descriptors1': d1, 'keypoints0': k0, 'keypoints1': k1,
            'image0': torch.empty(1, 3, 480, 640), 'image1': torch.empty(1, 3, 480, 640)}
    ref = O.gml_forward(sd, data, return_intermediate=True)
    out = net({k_: (v.to(dev) if k_.startswith(('desc', 'keyp')) else v) for k_, v in data.items()})
    assert torch.allclose(out['matching_scores0'].cpu(), ref['matching_scores0'], atol=2e-3)
    decisive = (ref['matching_scores0'] - 0.2).abs() > 5e-3
    assert torch.equal(out['matches0'].cpu()[decisive], ref['matches0'][decisive])
    with pytest.raises(ValueError):
        net({'descriptors0': d0.to(dev), 'descriptors1': d1.to(dev), 'keypoints0': k0.to(dev), 'keypoints1': k1.to(dev)})


def test_gml_both_sets_in_one_attention_launch(lib, dev):
    """m == n: self attention of both sets and both directions of the cross attention run as ONE launch each over
    [set 0 | set 1] (rotated key / value batches, pram_attention_tc_shift).  Must equal the per-set launches bit for bit, with
    and without per-pair keypoint counts, and stay inside the oracle tolerance."""
    from pram_b200.nets import _blocks as BL
    from pram_b200.nets.gml import GML
    sd = RL.random_gml_state(seed=7)
    net = GML({})
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).set_precision('bf16x3')
    g = torch.Generator().manual_seed(1)
    b, m = 3, 200
    d0 = torch.nn.functional.normalize(torch.randn(b, m, 128, generator=g), dim=-1)
    d1 = torch.nn.functional.normalize(torch.randn(b, m, 128, generator=g), dim=-1)
    d1[:, :120] = d0[:, 40:160] + 0.01 * torch.randn(b, 120, 128, generator=g)
    k0 = torch.rand(b, m, 2, generator=g) * torch.tensor([640., 480.])
    k1 = torch.rand(b, m, 2, generator=g) * torch.tensor([640., 480.])
    k1[:, :120] = k0[:, 40:160]
    data = {'descriptors0': d0.to(dev), 'descriptors1': d1.to(dev), 'keypoints0': k0.to(dev), 'keypoints1': k1.to(dev),
            'image_shape0': (1, 3, 640, 480), 'image_shape1': (1, 3, 640, 480)}
    saved = BL.MERGE_SETS
    try:
        for extra in ({}, {'num_keypoints0': torch.tensor([200, 150, 97]), 'num_keypoints1': torch.tensor([180, 200, 64])},
                      {'num_keypoints1': torch.tensor([33, 200, 199])}):
            res = {}
            for merge in (False, True):
                BL.MERGE_SETS = merge
                res[merge] = net({**data, **extra})
                torch.cuda.synchronize()
            for k_ in ('matches0', 'matches1', 'matching_scores0', 'matching_scores1'):
                assert torch.equal(res[False][k_], res[True][k_]), (k_, list(extra))
        ref = O.gml_forward(sd, {k_: (v.cpu() if torch.is_tensor(v) else v) for k_, v in data.items()})
        out = net(data)
        assert torch.allclose(out['matching_scores0'].cpu(), ref['matching_scores0'], atol=2e-3)
    finally:
        BL.MERGE_SETS = saved


def _adagml_case():
    g = torch.Generator().manual_seed(0)
    m = n = 400
    d0 = torch.nn.functional.normalize(torch.randn(1, m, 128, generator=g), dim=-1)
    perm = torch.randperm(n, generator=g)
    d1 = d0[:, perm] + 0.02 * torch.randn(1, n, 128, generator=g)
    k0 = torch.rand(1, m, 2, generator=g) * torch.tensor([640., 480.])
    return {'descriptors0': d0, 'descriptors1': d1, 'keypoints0': k0, 'keypoints1': k0[:, perm],
            'scores0': torch.rand(1, m, generator=g), 'scores1': torch.rand(1, n, generator=g),
            'image_shape0': (1, 3, 640, 480), 'image_shape1': (1, 3, 640, 480)}


@pytest.mark.parametrize('precision', PRECISIONS)
def test_adagml_pruning_vs_oracle(lib, dev, precision):
    """AdaGML with calibrated pooling weights: token pruning over several layers + early exit must follow
    the oracle's trace (data-dependent control flow), scores within tolerance."""
    from pram_b200.nets.adagml import AdaGML
    sd = RL.calibrated_adagml_state()
    data = _adagml_case()
    ref = O.adagml_forward(sd, data, return_trace=True)
    assert len(ref['trace']) >= 3 and ref['trace'][-1][1] < 200  # the case really prunes
    net = AdaGML({})
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).set_precision(precision)
    out = net({k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in data.items()})
    assert out['matches0'].shape == (1, 400) and out['matching_scores0'].shape == (1, 400)
    if precision == 'fp32':
        assert net.last_trace == ref['trace']
        assert torch.allclose(out['matching_scores0'].cpu(), ref['matching_scores0'], atol=2e-3)
        decisive = (ref['matching_scores0'] - 0.2).abs() > 5e-3
        assert torch.equal(out['matches0'].cpu()[decisive], ref['matches0'][decisive])
    else:  # confidences within ~1e-4 of a threshold may flip a token: allow a few tokens of slack per layer
        assert len(net.last_trace) == len(ref['trace'])
        for (l0, a0, b0), (l1, a1, b1) in zip(net.last_trace, ref['trace']):
            assert l0 == l1 and abs(a0 - a1) <= 4 and abs(b0 - b1) <= 4
    with pytest.raises(ValueError):
        net({k: (v.to(dev).repeat(2, 1, 1) if torch.is_tensor(v) and v.dim() == 3 else v) for k, v in data.items()})


def _adagml_case2(seed, m, n):
    g = torch.Generator().manual_seed(seed)
    d0 = torch.nn.functional.normalize(torch.randn(1, m, 128, generator=g), dim=-1)
    src = torch.randint(0, m, (n,), generator=g)
    d1 = d0[:, src] + 0.02 * torch.randn(1, n, 128, generator=g)
    k0 = torch.rand(1, m, 2, generator=g) * torch.tensor([640., 480.])
    return {'descriptors0': d0, 'descriptors1': d1, 'keypoints0': k0, 'keypoints1': k0[:, src],
            'scores0': torch.rand(1, m, generator=g), 'scores1': torch.rand(1, n, generator=g),
            'image_shape0': (1, 3, 640, 480), 'image_shape1': (1, 3, 640, 480)}


def _agree(out_i, out_s, ref, min_frac=0.97):
    """Device result vs the oracle for one pair: a confidence within ~1e-4 of a pruning threshold may flip a token, which
    perturbs its neighbours' scores slightly -- so almost all (not all) entries must agree."""
    decisive = (ref['matching_scores0'] - 0.2).abs() > 2e-2
    same = (out_i[decisive] == ref['matches0'][decisive]).float().mean().item()
    close = ((out_s - ref['matching_scores0']).abs() < 2e-2).float().mean().item()
    assert same >= min_frac and close >= min_frac, (same, close)


def _adagml_state(kind):
    """'seeded': random GML + calibrated pooling (prunes over layers 1-2, exits at layer 4, no confident matches);
    'shipped': the shipped GML weights + calibrated pooling (real matches; pairs exit at layer 1 or run all 9 layers)."""
    if kind == 'seeded':
        return RL.calibrated_adagml_state()
    g = RL.load_gml_state()
    if g is None:
        pytest.skip('GML checkpoint not staged')
    sd = RL.calibrated_adagml_state(gain=8.0, bias=0.15)
    sd.update(g)
    return sd


@pytest.mark.parametrize('kind', ['seeded', 'shipped'])
def test_adagml_device_path_vs_oracle_and_host_path(lib, dev, kind):
    """K17 on the device (csrc/adagml_ops.cu): pruning by stable compaction with per-pair token counts, stop flag + latched
    exit state, mean attention from the tcgen05 kernel -- against the oracle (pinned to the reference module) and against
    the reference-structured path of this repo (host decision per layer, boolean-mask gathers)."""
    from pram_b200.nets.adagml import AdaGML
    sd = _adagml_state(kind)
    data = _adagml_case()
    ref = O.adagml_forward(sd, data, return_trace=True)
    net = AdaGML({})
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).set_precision('bf16x3')
    gdata = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in data.items()}
    out = net.produce_matches_batched(gdata, check=True)
    tr = net.last_trace
    assert int(out['stop_layer'][0]) == ref['last_layer']
    assert len(tr) == len(ref['trace'])
    for (l0, a0, b0), (l1, a1, b1) in zip(tr, ref['trace']):
        assert l0 == l1 and abs(a0 - a1) <= 4 and abs(b0 - b1) <= 4
    assert int(out['num_tokens0'][0]) == tr[-1][1] and int(out['num_tokens1'][0]) == tr[-1][2]
    _agree(out['matches0'].cpu(), out['matching_scores0'].cpu(), ref)
    # pruned tokens come back unmatched with score exactly 0 like the reference's scatter (nets/adagml.py:389-394)
    gone = ref['matching_scores0'][0] == 0
    assert gone.sum() >= 400 - ref['trace'][-1][1]
    assert ((out['matching_scores0'][0].cpu() == 0) == gone).float().mean() > 0.98
    # the launch predicate (kernels of the layers after the exit return at once) is an optimisation only: the exit state was
    # latched, so running those layers for nothing gives the same bits
    net.config['device_early_exit'] = False
    full = net.produce_matches_batched(gdata, check=True)
    assert torch.equal(full['matches0'], out['matches0']) and torch.equal(full['matching_scores0'], out['matching_scores0'])
    net.config['device_pruning'] = False
    host = net(gdata)
    assert len(net.last_trace) == len(tr)
    _agree(out['matches0'].cpu(), out['matching_scores0'].cpu(),
           {'matches0': host['matches0'].cpu(), 'matching_scores0': host['matching_scores0'].cpu()})


def test_adagml_device_path_batched_pairs(lib, dev):
    """A batch of DIFFERENT pairs (one exits early, one runs on with heavy pruning, one too small to be pruned, padded to the
    common size) through the device path == each pair alone through the oracle: the reference cannot batch AdaGML at all
    (boolean-mask indexing, nets/adagml.py:358)."""
    from pram_b200.nets.adagml import AdaGML
    sd = _adagml_state('shipped')
    cases = [_adagml_case(), _adagml_case2(5, 400, 330), _adagml_case2(9, 200, 180), _adagml_case2(13, 300, 300)]
    refs = [O.adagml_forward(sd, c, return_trace=True) for c in cases]
    m = max(c['descriptors0'].shape[1] for c in cases)
    n = max(c['descriptors1'].shape[1] for c in cases)

    def pad(t, size):
        out = torch.zeros((1, size) + tuple(t.shape[2:]))
        out[:, :t.shape[1]] = t
        return out
    batch = {k: torch.cat([pad(c[k], m if k.endswith('0') else n) for c in cases]).to(dev)
             for k in ('descriptors0', 'descriptors1', 'keypoints0', 'keypoints1')}
    batch['image_shape0'] = batch['image_shape1'] = (1, 3, 640, 480)
    batch['num_keypoints0'] = torch.tensor([c['descriptors0'].shape[1] for c in cases], dtype=torch.int32)
    batch['num_keypoints1'] = torch.tensor([c['descriptors1'].shape[1] for c in cases], dtype=torch.int32)
    net = AdaGML({})
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).set_precision('bf16x3')
    out = net.produce_matches_batched(batch, check=True)
    stop = out['stop_layer'].cpu().tolist()
    trace = out['token_trace'].cpu()
    assert len(set(r['last_layer'] for r in refs)) > 1, 'the cases should exit at different layers'
    for i, (c, r) in enumerate(zip(cases, refs)):
        mi = c['descriptors0'].shape[1]
        assert stop[i] == r['last_layer'], (i, stop, [x['last_layer'] for x in refs])
        for (l, a, b_) in r['trace']:
            assert abs(int(trace[l, 0, i]) - a) <= 4 and abs(int(trace[l, 1, i]) - b_) <= 4
        _agree(out['matches0'][i, :mi].cpu()[None], out['matching_scores0'][i, :mi].cpu()[None], r)
        assert (out['matches0'][i, mi:] == -1).all() and (out['matching_scores0'][i, mi:] == 0).all()


def test_extract_sfd2_return_vs_oracle(lib, dev, golden):
    """Offline export variant (NMS radius 3, strict threshold, score-descending, x/(w/2)-1 sampling)."""
    if RL.weight_path(RL.SFD2_WEIGHT) is None:
        pytest.skip('SFD2 checkpoint not staged')
    from pram_b200.nets.sfd2 import ResNet4x, extract_sfd2_return
    sd = RL.load_sfd2_state()
    img = torch.from_numpy(O.polys_frame(120, 160, seed=3)).permute(2, 0, 1)[None]
    ref = O.sfd2_extract_return(sd, img, conf_th=0.005, topK=4096)
    net = ResNet4x()
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).set_precision('fp32')
    out = extract_sfd2_return(net, img, conf_th=0.005, topK=4096, scales=[1.0])
    assert out['keypoints'].dtype == np.float64 and out['descriptors'].shape[1] == 128
    ours = {(float(x), float(y)) for x, y in out['keypoints']}
    theirs = {(float(x), float(y)) for x, y in ref['keypoints']}
    assert len(ours & theirs) >= 0.97 * len(theirs)
    io = {k: i for i, k in enumerate(map(tuple, out['keypoints']))}
    ir = {k: i for i, k in enumerate(map(tuple, ref['keypoints']))}
    for k in list(ours & theirs)[:40]:
        assert np.abs(out['descriptors'][io[k]] - ref['descriptors'][ir[k]]).max() < 5e-4
        assert abs(out['scores'][io[k]] - ref['scores'][ir[k]]) < 2e-5
    assert np.all(np.diff(out['scores']) <= 1e-7)  # score-descending


def test_feature_matching_plugin(lib, dev, golden):
    """pose_estimator.feature_matching + the dynamic_load'ed matcher plugin (reference pose_estimator.py:45-86)."""
    if RL.weight_path(RL.GML_WEIGHT) is None:
        pytest.skip('GML checkpoint not staged')
    import pram_b200.localization.matchers as matchers
    from pram_b200.localization.base_model import dynamic_load
    from pram_b200.localization.pose_estimator import feature_matching
    g = golden('gml_selfmatch.npz')
    Model = dynamic_load(matchers, 'gml')
    model = Model({'name': 'gml', 'weight_path': str(RL.weight_path(RL.GML_WEIGHT)), 'sinkhorn_iterations': 20}).eval().to(dev)
    kp, desc, perm = g['keypoints0'], g['descriptors0'], g['perm']
    q = {'keypoints': kp, 'scores': np.ones(len(kp), np.float32), 'descriptors': desc, 'image_size': (160, 120)}
    ids = np.arange(len(perm)) + 1000
    ids[::5] = -1  # database keypoints without a 3-D point are excluded and the ids remapped
    db = {'keypoints': kp[perm], 'scores': np.ones(len(perm), np.float32), 'descriptors': desc[perm],
          'image_size': (160, 120), 'db_3D_ids': ids}
    m = feature_matching(q, db, model)
    ok = m >= 0
    assert ok.sum() > 50
    assert np.all(ids[m[ok]] != -1) and np.array_equal(perm[m[ok]], np.nonzero(ok)[0])


def test_sinkhorn_per_pair_sizes_vs_oracle(ops, dev):
    """Padded batch of Sinkhorn problems: pair b is the (m_b + 1) x (n_b + 1) problem of its first rows / columns.  Each
    must equal the oracle run on the unpadded block alone; padding comes back unmatched with zero scores."""
    g = torch.Generator().manual_seed(11)
    B_, M, N = 3, 200, 160
    dist = torch.randn(B_, M, N, generator=g) * 2
    mc, nc = [200, 131, 77], [160, 160, 50]
    for b in range(B_):
        for i in range(min(mc[b], nc[b]) // 2):
            dist[b, i, (i * 3) % nc[b]] += 15
    bs = torch.tensor(0.7)
    m0, m1, s0, s1, P = ops.sinkhorn_match(dist.to(dev), bs.to(dev), 20, 0.2, return_P=True,
                                           m_counts=torch.tensor(mc, dtype=torch.int32, device=dev),
                                           n_counts=torch.tensor(nc, dtype=torch.int32, device=dev))
    for b in range(B_):
        Pr = O.sinkhorn_with_dustbin(dist[b:b + 1, :mc[b], :nc[b]], bs, 20)
        i0, i1, t0, t1 = O.compute_matches(Pr, 0.2)
        Pb = P[b].cpu()
        assert torch.allclose(Pb[:mc[b], :nc[b]], Pr[0, :-1, :-1], rtol=5e-4, atol=1e-7)
        assert torch.allclose(Pb[M, :nc[b]], Pr[0, -1, :-1], rtol=5e-4, atol=1e-7) and torch.allclose(Pb[:mc[b], N], Pr[0, :-1, -1], rtol=5e-4, atol=1e-7)
        assert (Pb[mc[b]:M] == 0).all() and (Pb[:, nc[b]:N] == 0).all()
        dec = (t0[0] - 0.2).abs() > 1e-3
        assert torch.equal(m0[b, :mc[b]].cpu()[dec], i0[0][dec]) and torch.allclose(s0[b, :mc[b]].cpu(), t0[0], rtol=5e-4, atol=1e-6)
        assert (m0[b, mc[b]:] == -1).all() and (s0[b, mc[b]:] == 0).all() and (m1[b, nc[b]:] == -1).all()
        assert (i0[0] > -1).sum() > 10


@pytest.mark.skipif(RL.weight_path(RL.GML_WEIGHT) is None, reason='GML checkpoint not staged')
def test_padded_gml_and_segnetvit_equal_unpadded_runs(lib, dev, golden):
    """GML / SegNetViT on a padded batch with per-element keypoint counts == the same sets run alone without padding."""
    from pram_b200.nets.gml import GML
    from pram_b200.nets.segnetvit import SegNetViT
    g0 = golden('gml_selfmatch.npz')
    net = GML({})
    net.load_state_dict(RL.load_gml_state(), strict=True)
    net = net.to(dev)
    d = torch.from_numpy(g0['descriptors0'])
    k = torch.from_numpy(g0['keypoints0'])
    perm = torch.from_numpy(g0['perm'])
    n, n1 = d.shape[0], perm.shape[0]
    sizes = [(n, n1), (n - 17, n1 - 40)]
    Mp = Np = n + 9                                    # padded slot count
    D0, D1, K0, K1 = torch.zeros(2, Mp, 128), torch.zeros(2, Np, 128), torch.zeros(2, Mp, 2), torch.zeros(2, Np, 2)
    alone = []
    for b, (m_, n_) in enumerate(sizes):
        d0, k0, d1, k1 = d[:m_], k[:m_], d[perm][:n_], k[perm][:n_]
        D0[b, :m_], K0[b, :m_], D1[b, :n_], K1[b, :n_] = d0, k0, d1, k1
        alone.append(net({'descriptors0': d0[None].to(dev), 'keypoints0': k0[None].to(dev), 'descriptors1': d1[None].to(dev),
                          'keypoints1': k1[None].to(dev), 'image_shape0': (1, 3, 160, 120), 'image_shape1': (1, 3, 160, 120)}))
    out = net({'descriptors0': D0.to(dev), 'keypoints0': K0.to(dev), 'descriptors1': D1.to(dev), 'keypoints1': K1.to(dev),
               'num_keypoints0': torch.tensor([s[0] for s in sizes]), 'num_keypoints1': torch.tensor([s[1] for s in sizes]),
               'image_shape0': (1, 3, 160, 120), 'image_shape1': (1, 3, 160, 120)})
    for b, (m_, n_) in enumerate(sizes):
        a = alone[b]
        assert torch.allclose(out['matching_scores0'][b, :m_], a['matching_scores0'][0], atol=2e-5)
        dec = (a['matching_scores0'][0] - 0.2).abs() > 1e-3
        assert torch.equal(out['matches0'][b, :m_][dec], a['matches0'][0][dec]) and (a['matches0'][0] > -1).sum() > 30
        assert (out['matches0'][b, m_:] == -1).all() and (out['matches1'][b, n_:] == -1).all()
    with pytest.raises(ValueError):
        net({'descriptors0': D0.to(dev), 'keypoints0': K0.to(dev), 'descriptors1': D1.to(dev), 'keypoints1': K1.to(dev),
             'num_keypoints0': torch.tensor([3]), 'image_shape0': (1, 3, 160, 120), 'image_shape1': (1, 3, 160, 120)})
    # SegNetViT: logits of the real tokens are those of the unpadded run
    gs = golden('segnetvit_seed0.npz')
    sd = RL.random_segnetvit_state(int(gs['n_class']), seed=int(gs['seed']))
    vit = SegNetViT({'n_class': int(gs['n_class']), 'n_layers': 15, 'output_dim': 1024, 'descriptor_dim': 256})
    vit.load_state_dict(sd, strict=True)
    vit = vit.to(dev)
    shape = tuple(int(v) for v in gs['image_shape'])
    x = torch.from_numpy(gs['seg_descriptors'])
    kk = torch.from_numpy(gs['keypoints'])
    n = x.shape[0]
    X, KK = torch.zeros(2, n + 30, 256), torch.zeros(2, n + 30, 2)
    X[0, :n], KK[0, :n], X[1, :n - 25], KK[1, :n - 25] = x, kk, x[:n - 25], kk[:n - 25]
    img = torch.empty(shape, device='meta')
    pad = vit({'seg_descriptors': X.to(dev), 'keypoints': KK.to(dev), 'num_keypoints': torch.tensor([n, n - 25]), 'image': img})['prediction']
    a0 = vit({'seg_descriptors': x[None].to(dev), 'keypoints': kk[None].to(dev), 'image': img})['prediction']
    a1 = vit({'seg_descriptors': x[None, :n - 25].to(dev), 'keypoints': kk[None, :n - 25].to(dev), 'image': img})['prediction']
    assert torch.allclose(pad[0, :n], a0[0], atol=2e-5) and torch.allclose(pad[1, :n - 25], a1[0], atol=2e-5)
    assert np.abs(a0[0].cpu().numpy() - gs['prediction']).max() < 5e-3
