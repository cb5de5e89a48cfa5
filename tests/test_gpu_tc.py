"""tcgen05 tensor-core GEMM / implicit-GEMM convolution vs torch-CPU fp32 (the oracle's arithmetic).
split=3 (error-compensated bf16x3): tolerance 2e-4 of the output scale; split=1 (plain bf16): 2e-2."""
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = {1: 2e-2, 3: 2e-4}


@pytest.fixture(scope='module')
def ops(lib, dev):
    from pram_b200 import ops as _ops
    return _ops


def _relerr(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item()


@pytest.mark.parametrize('split', [1, 3])
@pytest.mark.parametrize('rows,k,n,bn', [(128, 64, 64, 64), (300, 256, 256, 0), (1000, 512, 768, 0), (77, 128, 113, 0),
                                         (4096, 256, 256, 128)])
def test_linear_tc(ops, dev, split, rows, k, n, bn):
    g = torch.Generator().manual_seed(rows + k + n)
    a = torch.randn(rows, k, generator=g)
    w = torch.randn(n, k, generator=g) / k ** 0.5
    bias = torch.randn(n, generator=g)
    res = torch.randn(rows, n, generator=g)
    ref = torch.relu(a @ w.t() + bias + res)
    A = ops.split_bf16(a.to(dev), split == 3)
    Wt = ops.split_bf16(w.to(dev), split == 3)
    out = torch.zeros(rows, n, device=dev)
    obf = ops.empty_split((rows, n), dev, split == 3) if n % 32 == 0 else None
    ops.linear_tc(A, k, rows, k, Wt, n, bias.to(dev), res.to(dev), n, True, out, n, obf, n, split=split, bn=bn)
    torch.cuda.synchronize()
    assert _relerr(out.cpu(), ref) < TOL[split]
    if obf is not None:
        assert _relerr(obf.float().cpu(), ref) < (TOL[split] if split == 3 else 3e-2)


@pytest.mark.parametrize('split', [1, 3])
@pytest.mark.parametrize('b,h,w,cin,cout', [(2, 24, 40, 64, 128), (1, 30, 40, 256, 256), (3, 17, 21, 128, 64)])
def test_conv3x3_s1_tc(ops, dev, split, b, h, w, cin, cout):
    g = torch.Generator().manual_seed(h * w + cin)
    x = torch.randn(b, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    bias = torch.randn(cout, generator=g)
    ref = torch.relu(torch.nn.functional.conv2d(x, wt, bias, padding=1))
    X = ops.split_bf16(x.permute(0, 2, 3, 1).contiguous().to(dev), split == 3)
    Wp = ops.split_bf16(wt.permute(2, 3, 0, 1).reshape(9, cout, cin).contiguous().to(dev), split == 3)
    out = ops.conv_tc(X, Wp, bias.to(dev), 3, 1, True, split, want_f32=True, want_bf=True, want_ps=True)
    torch.cuda.synchronize()
    assert _relerr(out['f32'].permute(0, 3, 1, 2).cpu(), ref) < TOL[split]
    assert _relerr(out['bf'].float().permute(0, 3, 1, 2).cpu(), ref) < max(TOL[split], 1e-2 if split == 1 else 0)
    # phase-split copy holds the same values: plane (y&1)*2+(x&1) at (y>>1, x>>1)
    ps = out['ps'].float().view(b, 2, 2, (h + 1) // 2, (w + 1) // 2, cout).cpu()
    full = out['bf'].float().cpu()
    for py in range(2):
        for px in range(2):
            sub = full[:, py::2, px::2]
            assert torch.equal(ps[:, py, px, :sub.shape[1], :sub.shape[2]], sub)


@pytest.mark.parametrize('split', [1, 3])
@pytest.mark.parametrize('b,h,w,cin,cout', [(2, 24, 40, 64, 64), (1, 31, 45, 128, 128)])
def test_conv3x3_s2_tc(ops, dev, split, b, h, w, cin, cout):
    """stride-2 convolution fed by the phase-split tensor a stride-1 conv produced."""
    g = torch.Generator().manual_seed(h + w + cin)
    x = torch.randn(b, cin, h, w, generator=g)
    w1 = torch.randn(cin, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    w2 = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    mid = torch.relu(torch.nn.functional.conv2d(x, w1, None, padding=1))
    ref = torch.nn.functional.conv2d(mid, w2, None, stride=2, padding=1)
    X = ops.split_bf16(x.permute(0, 2, 3, 1).contiguous().to(dev), split == 3)
    W1 = ops.split_bf16(w1.permute(2, 3, 0, 1).reshape(9, cin, cin).contiguous().to(dev), split == 3)
    W2 = ops.split_bf16(w2.permute(2, 3, 0, 1).reshape(9, cout, cin).contiguous().to(dev), split == 3)
    o1 = ops.conv_tc(X, W1, None, 3, 1, True, split, want_bf=False, want_ps=True)
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    o2 = ops.conv_tc(o1['ps'], W2, None, 3, 2, False, split, want_f32=True, want_bf=False, out_shape_hw=(ho, wo))
    torch.cuda.synchronize()
    assert o2['f32'].shape == (b, ho, wo, cout)
    assert _relerr(o2['f32'].permute(0, 3, 1, 2).cpu(), ref) < 2 * TOL[split]


def test_bmm_and_l2norm_tc(ops, dev):
    g = torch.Generator().manual_seed(0)
    bsz, m, n, k = 3, 200, 150, 256
    a = torch.randn(bsz, m, k, generator=g)
    b = torch.randn(bsz, n, k, generator=g)
    ref = torch.einsum('bmd,bnd->bmn', a, b)
    A, Bm = ops.split_bf16(a.to(dev)), ops.split_bf16(b.to(dev))
    out = torch.zeros(bsz, m, n, device=dev)
    ops.linear_tc(A, k, m, k, Bm, n, out_f32=out, ld_f32=n, split=3, batch=bsz, w_batched=True)
    torch.cuda.synchronize()
    assert _relerr(out.cpu(), ref) < 2e-4
    # fused row L2 normalisation (descriptor head)
    w = torch.randn(128, k, generator=g) / 16
    bias = torch.randn(128, generator=g) * 0.1
    ref = torch.nn.functional.normalize(a[0] @ w.t() + bias, dim=-1)
    out = torch.zeros(m, 128, device=dev)
    from pram_b200.ops import gemm_tc
    gemm_tc(ops.split_bf16(a[0].to(dev)), k, m, 1, 1, k, ops.split_bf16(w.to(dev)), 1, 1, 1, m, 128, [(0, 0, 0)],
            bias=bias.to(dev), out_f32=out, ld_f32=128, l2norm=True, split=3)
    torch.cuda.synchronize()
    assert _relerr(out.cpu(), ref) < 2e-4


@pytest.mark.parametrize('split', [3, 1])
@pytest.mark.parametrize('b,nq,nk', [(1, 128, 128), (2, 75, 130), (1, 1024, 1024), (2, 300, 1000)])
def test_attention_tc(ops, dev, split, b, nq, nk):
    """tcgen05 flash attention vs torch fp32 softmax(QK^T/8)V.  bf16x3: 2e-4 of the output scale."""
    g = torch.Generator().manual_seed(nq + nk)
    h = 4
    q, k, v = (torch.randn(b, h, n, 64, generator=g) for n in (nq, nk, nk))
    attn = torch.softmax(torch.einsum('bhid,bhjd->bhij', q, k) * 0.125, -1)
    ref = torch.einsum('bhij,bhjd->bhid', attn, v).transpose(1, 2).flatten(-2)  # [b, nq, 256]
    lo = split == 3
    nk_pad = (nk + 7) // 8 * 8
    vt = torch.zeros(b, h, 64, nk_pad)
    vt[..., :nk] = v.transpose(-1, -2)
    Q = ops.split_bf16(q.reshape(b * h, nq, 64).to(dev), lo)
    K = ops.split_bf16(k.reshape(b * h, nk, 64).to(dev), lo)
    VT = ops.split_bf16(vt.reshape(b * h, 64, nk_pad).to(dev), lo)
    out = torch.zeros(b, nq, 256, device=dev)
    obf = ops.empty_split((b, nq, 256), dev, lo)
    ops.attention_tc(Q, K, VT, b, h, nq, nk, nk_pad, 0.125, out, obf, 256, split)
    torch.cuda.synchronize()
    tol = 2e-4 if split == 3 else 2e-2
    assert _relerr(out.cpu(), ref) < tol
    assert _relerr(obf.float().cpu(), ref) < (tol if split == 3 else 3e-2)
    # V consumed directly as an MN-major operand (no transposition)
    V = ops.split_bf16(v.reshape(b * h, nk, 64).to(dev), lo)
    out = torch.zeros(b, nq, 256, device=dev)
    ops.attention_tc(Q, K, V, b, h, nq, nk, nk, 0.125, out, None, 256, split, v_mn=True)
    torch.cuda.synchronize()
    assert _relerr(out.cpu(), ref) < tol, 'MN-major V operand'


def test_attention_prep_matches_rotary_split(ops, dev):
    g = torch.Generator().manual_seed(0)
    b, n, h = 2, 70, 4
    qkv = torch.randn(b * n, 768, generator=g).to(dev)
    cos = torch.rand(b * n, 32, generator=g).to(dev)
    sin = torch.rand(b * n, 32, generator=g).to(dev)
    q = torch.empty(b * n, 256, device=dev); k = torch.empty_like(q); v = torch.empty_like(q)
    ops.rotary_split(qkv, 3, b, n, h, cos, sin, 0.5, q, k, v)
    Q, K, VT, n_pad = ops.attention_prep(qkv, 3, b, n, h, cos, sin, 0.5, 3)
    torch.cuda.synchronize()
    assert torch.allclose(Q.float().view(-1), q.view(-1), atol=1e-5, rtol=1e-5)
    assert torch.allclose(K.float().view(-1), k.view(-1), atol=1e-5, rtol=1e-5)
    vref = v.view(b, h, n, 64).transpose(-1, -2)
    assert torch.allclose(VT.float().view(b, h, 64, n_pad)[..., :n], vref, atol=1e-5, rtol=1e-5)
    assert (VT.float().view(b, h, 64, n_pad)[..., n:] == 0).all()


@pytest.mark.parametrize('split', [1, 3])
@pytest.mark.parametrize('b,h,w', [(2, 16, 64), (1, 30, 40), (3, 9, 33), (1, 120, 160)])
def test_gconv3x3_tc(ops, dev, split, b, h, w):
    """32-group 3x3 convolution on warp-level bf16 MMAs (ResBlock.conv2, reference nets/sfd2.py:100-124) vs
    torch-CPU fp32 grouped conv2d; ragged tiles (h % 8, w % 32 != 0) exercise the zero-filled halo."""
    c, groups = 256, 32
    g = torch.Generator().manual_seed(h * w + split)
    x = torch.randn(b, c, h, w, generator=g)
    wt = torch.randn(c, 8, 3, 3, generator=g) / 72 ** 0.5
    bias = torch.randn(c, generator=g)
    ref = torch.relu(torch.nn.functional.conv2d(x, wt, bias, padding=1, groups=groups))
    # same packing as ResNet4x.prepare(): [tap][ci][co][group]
    wp = wt.view(groups, 8, 8, 3, 3).permute(3, 4, 2, 1, 0).reshape(9, 8, 8, groups).contiguous().to(dev)
    X = ops.split_bf16(x.permute(0, 2, 3, 1).contiguous().to(dev), split == 3)
    out = ops.gconv3x3_tc(X, wp, bias.to(dev), True, split)
    torch.cuda.synchronize()
    assert _relerr(out.float().permute(0, 3, 1, 2).cpu(), ref) < (TOL[split] if split == 3 else 3e-2)
    # and against the fp32 CUDA-core kernel on the same (quantised) input
    simt = ops.gconv3x3_f32(X.float(), wp, bias.to(dev), True)
    assert _relerr(out.float(), simt) < (TOL[split] if split == 3 else 3e-2)


@pytest.mark.parametrize('case', ['conv', 'linear', 'linear_odd', 'qkv'])
def test_gemm_cluster_multicast_bit_identical(ops, dev, case):
    """2-CTA clusters with the weight tile TMA-multicast to both CTAs must give bit-identical results to the
    single-CTA schedule (same MMA order per tile), including an odd number of M tiles (one CTA of the last pair idles
    on an all-zero, out-of-bounds A tile) and the fused qkv epilogue."""
    g = torch.Generator().manual_seed(7)
    outs = {}
    for cl in (1, 2):
        ops.GEMM_CLUSTER = cl
        try:
            if case == 'conv':
                x = torch.randn(2, 40, 48, 128, generator=g.manual_seed(1))
                wt = torch.randn(9, 256, 128, generator=g) * 0.03
                X, Wp = ops.split_bf16(x.to(dev), True), ops.split_bf16(wt.to(dev), True)
                o = ops.conv_tc(X, Wp, torch.ones(256, device=dev), 3, 1, True, 3, want_f32=True, want_ps=True)
                outs[cl] = [o['f32'], o['bf'].hi, o['bf'].lo, o['ps'].hi]
            elif case in ('linear', 'linear_odd'):
                rows = 4096 if case == 'linear' else 128 * 37 + 5  # 38 M tiles (even) / ragged last tile + odd handling
                rows = rows if case == 'linear' else 128 * 36 + 5   # 37 M tiles: odd
                a = torch.randn(rows, 512, generator=g.manual_seed(2))
                w = torch.randn(384, 512, generator=g) * 0.05
                res = torch.randn(rows, 384, generator=g)
                A, Wt = ops.split_bf16(a.to(dev), True), ops.split_bf16(w.to(dev), True)
                of = torch.zeros(rows, 384, device=dev)
                ob = ops.empty_split((rows, 384), dev, True, zero=True)
                ops.linear_tc(A, 512, rows, 512, Wt, 384, torch.ones(384, device=dev), res.to(dev), 384, False, of, 384, ob, 384, split=3)
                outs[cl] = [of, ob.hi, ob.lo]
            else:
                b, n = 3, 700
                T = b * n
                a = torch.randn(T, 256, generator=g.manual_seed(3))
                w = torch.randn(768, 256, generator=g) * 0.05
                cos, sin = torch.rand(T, 32, generator=g), torch.rand(T, 32, generator=g)
                A, Wt = ops.split_bf16(a.to(dev), True), ops.split_bf16(w.to(dev), True)
                q, k, v = (ops.empty_split((T, 256), dev, True, zero=True) for _ in range(3))
                qkv = {'mode': 1, 'scale': 1.0, 'cos': cos.to(dev), 'sin': sin.to(dev), 'q': q, 'k': k, 'v': v,
                       'seg_split': T, 'seg_n0': n, 'seg_n1': n}
                ops.linear_tc(A, 256, T, 256, Wt, 768, torch.ones(768, device=dev), split=3, bn=256, qkv=qkv)
                outs[cl] = [q.hi, q.lo, k.hi, k.lo, v.hi, v.lo]
            torch.cuda.synchronize()
        finally:
            ops.GEMM_CLUSTER = 0
    for t1, t2 in zip(outs[1], outs[2]):
        assert torch.equal(t1, t2)


@pytest.mark.parametrize('split', [3, 1])
@pytest.mark.parametrize('b,h,w', [(2, 16, 128), (1, 31, 77), (3, 24, 300)])
def test_conv1a_tc(ops, dev, split, b, h, w):
    """conv1a on tcgen05 (CTA-built im2col operand) vs torch-CPU fp32 conv2d + ReLU, through the phase-split layout,
    and vs the CUDA-core kernel; ragged widths / odd heights exercise the zero rows and the phase-split padding."""
    g = torch.Generator().manual_seed(h * w)
    x = torch.rand(b, 3, h, w, generator=g) * 2 - 1
    wt = torch.randn(64, 3, 3, 3, generator=g) * 0.2
    bias = torch.randn(64, generator=g) * 0.1
    ref = torch.relu(torch.nn.functional.conv2d(x, wt, bias, padding=1)).permute(0, 2, 3, 1)  # NHWC
    wp = wt.permute(2, 3, 1, 0).reshape(27, 64).contiguous().to(dev)  # k = (ry*3 + rx)*3 + c
    outs = {}
    for tc in (True, False):
        ops.CONV1A_TC = tc
        try:
            ps, _ = ops.conv1a(x.to(dev), wp, bias.to(dev), split)
            torch.cuda.synchronize()
        finally:
            ops.CONV1A_TC = True
        hp, wq = (h + 1) // 2, (w + 1) // 2
        full = torch.zeros(b, hp * 2, wq * 2, 64)
        psf = ps.float().view(b, 2, 2, hp, wq, 64).cpu()
        for py in range(2):
            for px in range(2):
                full[:, py::2, px::2] = psf[:, py, px]
        outs[tc] = full[:, :h, :w]
    tol = TOL[split] if split == 3 else 3e-2
    assert _relerr(outs[False], ref) < tol
    assert _relerr(outs[True], ref) < tol


@pytest.mark.parametrize('p16', [False, True])
@pytest.mark.parametrize('kv_tile', [64, 128])
@pytest.mark.parametrize('b,nq,nk', [(1, 128, 128), (2, 75, 130), (1, 1024, 1024), (2, 300, 1000), (1, 4096, 1024)])
def test_attention_tc_kv_tile_variants(ops, dev, kv_tile, b, nq, nk, p16):
    """Both key-tile variants of the flash-attention kernel (64: two CTAs per SM, 128: one) vs torch fp32, with the
    probabilities fed back as bf16 hi / lo planes (V bf16 planes) and as one fp16 plane (V fp16 planes, ``p16``)."""
    g = torch.Generator().manual_seed(nq + nk + kv_tile)
    h = 4
    q, k, v = (torch.randn(b, h, n, 64, generator=g) for n in (nq, nk, nk))
    attn = torch.softmax(torch.einsum('bhid,bhjd->bhij', q, k) * 0.125, -1)
    ref = torch.einsum('bhij,bhjd->bhid', attn, v).transpose(1, 2).flatten(-2)
    Q = ops.split_bf16(q.reshape(b * h, nq, 64).to(dev), True)
    K = ops.split_bf16(k.reshape(b * h, nk, 64).to(dev), True)
    splitv = ops.split_f16 if p16 else ops.split_bf16
    V = splitv(v.reshape(b * h, nk, 64).to(dev), True)
    nk_pad = (nk + 7) // 8 * 8
    vt = torch.zeros(b, h, 64, nk_pad)
    vt[..., :nk] = v.transpose(-1, -2)
    VT = splitv(vt.reshape(b * h, 64, nk_pad).to(dev), True)
    ops.ATT_KV_TILE = kv_tile
    try:
        out = torch.zeros(b, nq, 256, device=dev)
        obf = ops.empty_split((b, nq, 256), dev, True)
        ops.attention_tc(Q, K, V, b, h, nq, nk, nk, 0.125, out, obf, 256, 3, v_mn=True, v_f16=p16)
        out2 = torch.zeros(b, nq, 256, device=dev)
        ops.attention_tc(Q, K, VT, b, h, nq, nk, nk_pad, 0.125, out2, None, 256, 3, v_f16=p16)
        torch.cuda.synchronize()
    finally:
        ops.ATT_KV_TILE = 0
    tol = 6e-4 if p16 else 2e-4   # one fp16 probability plane: softmax weights exact to 2^-12 relative (bf16 hi / lo: 2^-17)
    assert _relerr(out.cpu(), ref) < tol
    assert _relerr(obf.float().cpu(), ref) < tol
    assert _relerr(out2.cpu(), ref) < tol, 'V^T (K-major) operand path'


@pytest.mark.parametrize('split', [1, 3])
@pytest.mark.parametrize('rows', [128, 77, 1000, 148 * 128 * 2 + 300])
def test_mlp_block_tc(ops, dev, split, rows):
    """Fused block tail (csrc/mlp_block_tc.cu) vs the unfused torch-CPU arithmetic of the reference block
    (nets/segnetvit.py:104-106): proj -> cat -> Linear -> LayerNorm -> GELU -> Linear -> + x.  Ragged row counts
    exercise the zero-filled tail tile; the largest case gives every CTA several tiles (barrier phases across tiles,
    accumulator hand-over between tiles)."""
    import torch.nn.functional as F
    from pram_b200.nets import _blocks as B
    g = torch.Generator().manual_seed(rows)
    proj = torch.nn.Linear(256, 256)
    mlp = B.mlp_holder(512, 512, 256)
    with torch.no_grad():
        for prm in list(proj.parameters()) + list(mlp.parameters()):
            prm.copy_(torch.randn(prm.shape, generator=g) * (0.05 if prm.dim() == 2 else 0.5))
        mlp[1].weight.add_(1.0)
    x = torch.randn(rows, 256, generator=g)
    ctx = torch.randn(rows, 256, generator=g)
    with torch.no_grad():
        ref = x + mlp(torch.cat([x, proj(ctx)], -1))
    pk = B.pack_block_tail(proj.to(dev), mlp.to(dev))
    lo = split == 3
    if not lo:
        pk['blk.w1.tc'] = ops.Split(pk['blk.w1.tc'].hi, None)
        pk['blk.w3.tc'] = ops.Split(pk['blk.w3.tc'].hi, None)
    cat = torch.cat([x, ctx], 1).to(dev)
    cat_bf = ops.split_bf16(cat, lo)
    w3 = pk['blk.w3.tc']
    out = torch.full((rows, 512), 7.0, device=dev)
    obf = ops.empty_split((rows, 512), dev, lo, zero=True)
    # (a) fp32 residual rows given, fp32 + split-bf16 outputs
    ops.mlp_block_tc(cat_bf, 512, rows, pk['blk.w1.tc'], w3, pk['blk.tables'], cat, 512, out, 512, obf, 512, split=split)
    torch.cuda.synchronize()
    tol = {1: 3e-2, 3: 2e-4}[split]
    assert _relerr(out[:, :256].cpu(), ref) < tol
    assert _relerr(obf.float()[:, :256].cpu(), ref) < (tol if lo else 4e-2)
    # the right halves of the output rows (the next block's context slots) are not touched
    assert (out[:, 256:] == 7.0).all() and (obf.hi[:, 256:] == 0).all()
    # (b) the production form: residual taken from the bf16 planes of x, split-bf16 output only; repeated launches on the
    # same buffers agree bit for bit (no state left behind in TMEM / barriers between launches)
    o1 = ops.empty_split((rows, 512), dev, lo, zero=True)
    o2 = ops.empty_split((rows, 512), dev, lo, zero=True)
    for o in (o1, o2):
        ops.mlp_block_tc(cat_bf, 512, rows, pk['blk.w1.tc'], w3, pk['blk.tables'], None, 0, None, 0, o, 512, split=split)
    torch.cuda.synchronize()
    assert torch.equal(o1.hi, o2.hi) and (not lo or torch.equal(o1.lo, o2.lo))
    ref_b = (ref - x) + cat_bf.float()[:, :256].cpu()  # residual = hi + lo of x (exact in split 3 up to 2^-17)
    assert _relerr(o1.float()[:, :256].cpu(), ref_b) < (tol if lo else 4e-2)


@pytest.mark.parametrize('kv_tile', [64, 128])
@pytest.mark.parametrize('split', [3, 1])
@pytest.mark.parametrize('b,nq,nk,counts', [(1, 128, 128, False), (2, 75, 130, False), (2, 300, 1000, True), (1, 1024, 400, True),
                                            (3, 400, 400, True)])
def test_attention_colmean_tc(ops, dev, split, kv_tile, b, nq, nk, counts):
    """AdaGML's mean attention per key (nets/adagml.py:148: mean over heads, then over queries) on the tensor cores: row
    statistics from the flash kernel + column sums from S^T tiles, vs torch fp32 -- with ragged sizes and with per-batch
    counts of valid queries / keys (padding takes no part on either side)."""
    g = torch.Generator().manual_seed(nq * 3 + nk + kv_tile)
    h = 4
    q, k, v = (torch.randn(b, h, n, 64, generator=g) for n in (nq, nk, nk))
    qc = torch.tensor([nq - 17 * i for i in range(b)], dtype=torch.int32) if counts else None
    kc = torch.tensor([nk - 29 * i - 3 for i in range(b)], dtype=torch.int32) if counts else None
    ref = torch.zeros(b, nk)
    ctx_ref = torch.zeros(b, nq, 256)
    for i in range(b):
        nqi, nki = (int(qc[i]), int(kc[i])) if counts else (nq, nk)
        attn = torch.softmax(torch.einsum('hid,hjd->hij', q[i, :, :nqi], k[i, :, :nki]) * 0.125, -1)
        ref[i, :nki] = attn.mean(0).mean(0)
        ctx_ref[i, :nqi] = torch.einsum('hij,hjd->hid', attn, v[i, :, :nki]).transpose(0, 1).flatten(-2)
    lo = split == 3
    Q = ops.split_bf16(q.reshape(b * h, nq, 64).to(dev), lo)
    K = ops.split_bf16(k.reshape(b * h, nk, 64).to(dev), lo)
    p16 = bool(b & 1)  # probabilities as one fp16 plane (V as fp16 planes) on some of the cases
    V = (ops.split_f16 if p16 else ops.split_bf16)(v.reshape(b * h, nk, 64).to(dev), lo)
    ops.ATT_KV_TILE = kv_tile
    try:
        out = torch.zeros(b, nq, 256, device=dev)
        lse = torch.full((b * h, ops.lse_ld(nq)), float('nan'), device=dev)
        colsum = torch.full((b * h, ops.lse_ld(nk)), float('nan'), device=dev)
        cm = torch.full((b * nk, 2), float('nan'), device=dev)
        qcd, kcd = (qc.to(dev), kc.to(dev)) if counts else (None, None)
        ops.attention_tc(Q, K, V, b, h, nq, nk, nk, 0.125, out, None, 256, split, v_mn=True, nk_counts=kcd, lse_out=lse,
                         v_f16=p16)
        ops.attention_colmean_tc(K, Q, b, h, nk, nq, 0.125, lse, colsum, cm[:, 1], 2, split, nq_counts=qcd)
        torch.cuda.synchronize()
    finally:
        ops.ATT_KV_TILE = 0
    got = cm[:, 1].view(b, nk).cpu()
    tol = (6e-4 if p16 else 2e-4) if split == 3 else 2e-2
    for i in range(b):
        nqi, nki = (int(qc[i]), int(kc[i])) if counts else (nq, nk)
        assert _relerr(out[i, :nqi].cpu(), ctx_ref[i, :nqi]) < tol
        assert _relerr(got[i, :nki], ref[i, :nki]) < (5e-4 if split == 3 else 3e-2)
        assert abs(float(got[i, :nki].sum()) - 1.0) < (1e-3 if split == 3 else 2e-2)  # every query row sums to 1
    assert torch.isnan(cm[:, 0]).all()  # the other column of the [T, 2] pooling input is untouched


def test_conv1x1_residual_planes_fp16_copy_and_l2norm(ops, dev):
    """ResBlock tail as the mixed mode runs it: 1x1 conv + residual from split-bf16 planes + ReLU, with the fp32 map, the
    phase-split planes and ONE fp16 plane written by the same epilogue -- the fp16 plane must equal the separate cast of the
    fp32 output bit for bit; then the L2-normalised descriptor projection (vector-bias row norms) on that plane."""
    g = torch.Generator().manual_seed(5)
    b, h, w, c = 2, 30, 40, 256
    x = torch.randn(b, h, w, c, generator=g)
    r = torch.randn(b, h, w, c, generator=g)
    wt = torch.randn(1, c, c, generator=g) / c ** 0.5
    bias = torch.randn(c, generator=g)
    X, R, Wp = ops.split_bf16(x.to(dev)), ops.split_bf16(r.to(dev)), ops.split_bf16(wt.to(dev))
    out = ops.conv_tc(X, Wp, bias.to(dev), 1, 1, True, 3, res_bf=R, want_f32=True, want_bf=False, want_ps=True, want_h16=True)
    both = ops.conv_tc(X, Wp, bias.to(dev), 1, 1, True, 3, res_bf=R, want_f32=True, want_bf=True)
    torch.cuda.synchronize()
    ref = torch.relu(X.float().cpu() @ wt[0].t() + bias + R.float().cpu())
    assert _relerr(out['f32'].cpu(), ref) < TOL[3]
    assert torch.equal(out['f32'], both['f32'])
    assert torch.equal(out['h16'].hi.view(torch.float16), ops.as_f16_plane(out['f32']).hi.view(torch.float16))
    assert 'bf' not in out
    # descriptor projection: fp16 plane in, fp32 L2-normalised rows out (N = 128 <= BN)
    wd = (torch.randn(1, 128, c, generator=g) / c ** 0.5)
    bd = torch.randn(128, generator=g)
    Wd = ops.Split(wd.half().to(dev).view(torch.bfloat16), None)
    d = ops.conv_tc(out['h16'], Wd, bd.to(dev), 1, 1, False, 1, want_f32=True, want_bf=False, l2norm=True, f16=True)['f32']
    torch.cuda.synchronize()
    a16 = out['h16'].hi.view(torch.float16).float().cpu()
    dref = torch.nn.functional.normalize(a16 @ wd.half().float()[0].t() + bd, dim=-1)
    assert (d.cpu() - dref).abs().max().item() < 2e-5
    assert (d.cpu().norm(dim=-1) - 1).abs().max().item() < 1e-5
