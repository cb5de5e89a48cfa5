"""Per-kernel roofline table: every named kernel of the hot path (SURVEY.md 8a / DESIGN.md section 5) launched in
isolation at the 7Scenes bench shape (B frames of 640x480, K=1024), timed with CUDA events on the launch stream,
L2 flushed between repetitions.  Algorithmic bytes / FLOPs per launch are the SURVEY.md 8d per-unit figures times
the units one launch processes.  Run under gpurun:

    python tools/bench_kernels.py [B] [out.json]

Prints one JSON object per kernel and a markdown table; `ncu` can wrap the same command (see profiles/README.md).
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pram_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
OUT = sys.argv[2] if len(sys.argv) > 2 else None
H, W, K = 480, 640, 1024
dev = torch.device('cuda')
torch.manual_seed(0)
peaks = json.loads((ROOT / 'MEASURED_PEAKS.json').read_text()) if (ROOT / 'MEASURED_PEAKS.json').exists() else {}
HBM = float(peaks.get('hbm_gbs', 6650.0))
TC = float(peaks.get('bf16_tflops', 1590.0))
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def timeit(fn, reps=7, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


rows = []


def report(name, ms, nbytes=None, flops=None, bound='hbm', note=''):
    r = {'kernel': name, 'ms': ms, 'bound': bound, 'note': note, 'frames': B}
    if nbytes is not None:
        r['algorithmic_MB'] = nbytes / 1e6
        r['GBps'] = nbytes / (ms * 1e-3) / 1e9
        if bound == 'hbm':
            r['frac'] = r['GBps'] / HBM
    if flops is not None:
        r['algorithmic_GFLOP'] = flops / 1e9
        r['TFLOPs'] = flops / (ms * 1e-3) / 1e12
        if bound == 'tensor':
            r['frac'] = r['TFLOPs'] / TC
    rows.append(r)
    print(json.dumps(r), flush=True)


# ---- conv1a ------------------------------------------------------------------------------------------
img = torch.rand(B, 3, H, W, device=dev)
w1 = torch.randn(27, 64, device=dev) * 0.1
b1 = torch.randn(64, device=dev) * 0.1
report('conv1a_kernel (3->64, NCHW f32 -> phase-split bf16x2)', timeit(lambda: ops.conv1a(img, w1, b1, 3)),
       nbytes=B * (4 * 3 * H * W + 2 * 2 * 64 * H * W), flops=B * 2 * 27 * 64 * H * W)

# ---- tcgen05 implicit GEMM at the conv-stack shapes --------------------------------------------------------
def conv_case(name, h, w, cin, cout, ks, stride=1):
    x = ops.split_bf16(torch.randn(B, h, w, cin, device=dev), True)
    wt = ops.split_bf16(torch.randn(ks * ks, cout, cin, device=dev) * 0.02, True)
    bias = torch.randn(cout, device=dev)
    ms = timeit(lambda: ops.conv_tc(x, wt, bias, ks, 1, True, 3))
    fl = 2.0 * B * h * w * cout * cin * ks * ks
    by = B * h * w * (cin + cout) * 4 + ks * ks * cin * cout * 4
    report(f'gemm_tc_kernel {name}', ms, nbytes=by, flops=fl, bound='tensor',
           note='bf16x3: 3 tensor FLOPs issued per algorithmic FLOP (frac <= 1/3)')


conv_case('conv3x3 256->256 @120x160 (conv3b)', 120, 160, 256, 256, 3)
conv_case('conv3x3 128->128 @240x320 (conv2a-like)', 240, 320, 128, 128, 3)
conv_case('conv1x1 256->256 @120x160 (conv4.x.conv1/3)', 120, 160, 256, 256, 1)

xr_ = ops.split_bf16(torch.randn(B, 120, 160, 256, device=dev), True)
rr_ = ops.split_bf16(torch.randn(B, 120, 160, 256, device=dev), True)
wr_ = ops.split_bf16(torch.randn(1, 256, 256, device=dev) * 0.05, True)
br_ = torch.randn(256, device=dev)
report('gemm_tc_kernel conv1x1 256->256 + residual planes + ReLU @120x160 (conv4.x.conv3)',
       timeit(lambda: ops.conv_tc(xr_, wr_, br_, 1, 1, True, 3, res_bf=rr_)), nbytes=B * 120 * 160 * 256 * 12, bound='hbm',
       note='reads A hi/lo + residual hi/lo, writes hi/lo: 12 B per output element')
x16_ = ops.as_f16_plane(torch.randn(B, 120, 160, 256, device=dev))
wd_ = ops.Split((torch.randn(1, 128, 256, device=dev) * 0.05).half().view(torch.bfloat16), None)
bd_ = torch.randn(128, device=dev)
report('gemm_tc_kernel conv1x1 256->128 fp16 + L2 norm @120x160 (convDb)',
       timeit(lambda: ops.conv_tc(x16_, wd_, bd_, 1, 1, False, 1, want_f32=True, want_bf=False, l2norm=True, f16=True)),
       nbytes=B * 120 * 160 * (256 * 2 + 128 * 4), bound='hbm')
del xr_, rr_, x16_

# ---- linear layers at transformer shapes -------------------------------------------------------------------
def lin_case(name, rows_, k, n):
    a = ops.split_bf16(torch.randn(rows_, k, device=dev), True)
    wt = ops.split_bf16(torch.randn(n, k, device=dev) * 0.05, True)
    bias = torch.randn(n, device=dev)
    of = torch.empty(rows_, n, device=dev)
    ms = timeit(lambda: ops.linear_tc(a, k, rows_, k, wt, n, bias, out_f32=of, ld_f32=n, split=3))
    report(f'gemm_tc_kernel {name}', ms, nbytes=rows_ * (k * 4 + n * 4) + n * k * 4, flops=2.0 * rows_ * k * n, bound='tensor',
           note='bf16x3; fp32 output')


lin_case(f'Linear 512->512, {B * K} tokens (mlp.0)', B * K, 512, 512)
lin_case(f'Linear 256->256, {B * K} tokens (proj)', B * K, 256, 256)

# ---- grouped conv -------------------------------------------------------------------------------------------
xg = torch.randn(B, 120, 160, 256, device=dev)
wg = torch.randn(9 * 8 * 8 * 32, device=dev) * 0.1
bg = torch.randn(256, device=dev)
report('gconv3x3_kernel (32 groups x 8, 256 ch @120x160)', timeit(lambda: ops.gconv3x3_split(xg, wg, bg, True, 3)),
       nbytes=B * 120 * 160 * 256 * (4 + 4), flops=B * 2.0 * 9 * 8 * 256 * 120 * 160)

xgs = ops.split_bf16(xg, True)
report('gconv_mma_kernel (32 groups x 8, 256 ch @120x160, bf16x3 mma.sync)', timeit(lambda: ops.gconv3x3_tc(xgs, wg, bg, True, 3)),
       nbytes=B * 120 * 160 * 256 * (4 + 4), flops=B * 2.0 * 9 * 8 * 256 * 120 * 160)

# ---- score map, NMS, selection --------------------------------------------------------------------------------
logits = torch.randn(B, H // 8, W // 8, 65, device=dev) * 3
report('score_map_kernel', timeit(lambda: ops.score_map(logits)), nbytes=B * 4 * (65 * H * W // 64 + H * W))
score = ops.score_map(logits)
from pram_b200._lib import call, ptr, stream_ptr  # noqa: E402
cap = 16384
cand = torch.empty((B, cap), device=dev, dtype=torch.int64)
counts = torch.empty((2, B), device=dev, dtype=torch.int32)


def nms_only():
    call('pram_nms_candidates', ptr(score), B, H, W, 4, 0.0025, 0.005, None, ptr(cand), cap, ptr(counts[0]), ptr(counts[1]),
         stream_ptr())


report('nms_kernel (r=4, 2 rounds, candidate emission)', timeit(nms_only), nbytes=B * 4 * H * W,
       note='reads the score map once; candidates are O(K) bytes')
kp = torch.empty((B, K, 2), device=dev)
sc = torch.empty((B, K), device=dev)
nn_ = torch.empty((B,), device=dev, dtype=torch.int32)


def select_only():
    call('pram_select_keypoints', ptr(cand), cap, ptr(counts[0]), ptr(counts[1]), ptr(score), B, H, W, 0.0025, 0.005, 128, K, 4,
         0, 0, ptr(kp), ptr(sc), ptr(nn_), K, None, stream_ptr())


nms_only()
torch.cuda.synchronize()
ncand = int(counts[0].float().mean().item())
report('select_kernel + fill_scores_kernel (radix select top-K + sort)', timeit(select_only), nbytes=B * (8 * ncand + 12 * K),
       bound='latency', note=f'{ncand} candidates/frame; one CTA per frame')

# ---- sampling / positional encoding --------------------------------------------------------------------------
select_only()
dmap = torch.randn(B, 120, 160, 128, device=dev)
mid = torch.randn(B, 120, 160, 256, device=dev)
report('sample_kernel (descriptors 128 ch, L2 norm)', timeit(lambda: ops.sample_features(dmap, kp, nn_, 4, True)),
       nbytes=B * K * 128 * (4 * 4 + 4), note='gather: 4 taps x 512 B lines per keypoint')
report('sample_kernel (mid features 256 ch)', timeit(lambda: ops.sample_features(mid, kp, nn_, 4, False)),
       nbytes=B * K * 256 * (4 * 4 + 4))
wr = torch.randn(32, 2, device=dev)
report('posenc_kernel', timeit(lambda: ops.posenc(kp, W, H, wr)), nbytes=B * K * (8 + 2 * 32 * 4), bound='latency')

lg = torch.randn(B, K, 113, device=dev)
report('top_classes_kernel + rank_entries_kernel (process_segmentations, 113 classes)', timeit(lambda: ops.rank_landmarks(lg, None, 20, 8)),
       nbytes=B * K * 113 * 4, bound='latency')

# ---- LayerNorm + GELU -------------------------------------------------------------------------------------------
xl = torch.randn(B * K, 512, device=dev)
gl, bl = torch.randn(512, device=dev), torch.randn(512, device=dev)
ol = ops.empty_split((B * K, 512), dev, True)
report('layernorm_gelu_vec_kernel (512 ch, f32 -> bf16x2)', timeit(lambda: ops.layernorm_gelu_split(xl, gl, bl, 512, ol)),
       nbytes=B * K * 512 * (4 + 4))

# ---- attention ---------------------------------------------------------------------------------------------------
q = ops.split_bf16(torch.randn(B * 4, K, 64, device=dev), True)
k_ = ops.split_bf16(torch.randn(B * 4, K, 64, device=dev), True)
v = ops.split_bf16(torch.randn(B * 4, K, 64, device=dev), True)
ctx = ops.empty_split((B * K, 256), dev, True)
report('attention_tc_kernel (4 heads, N=1024)', timeit(lambda: ops.attention_tc(q, k_, v, B, 4, K, K, K, 0.125, None, ctx, 256, 3, v_mn=True)),
       nbytes=B * K * 256 * 4 * 4, flops=4.0 * B * 4 * K * K * 64, bound='tensor', note='bf16x3 QK^T and PV')

# ---- Sinkhorn + match extraction ------------------------------------------------------------------------------------
dist = torch.randn(B, K, K, device=dev) * 3
bin_score = torch.tensor(1.0, device=dev)
for g in (8, 4, 2):
    report(f'sinkhorn_match_kernel (20 it, 1024x1024, cluster {g})', timeit(lambda: ops.sinkhorn_match(dist, bin_score, 20, 0.2, cluster=g)),
           nbytes=B * 46 * 4 * (K + 1) * (K + 1), note='SURVEY 8d streaming model 46 x 4 x (M+1)(N+1) B; resident/L2 design moves 23 sweeps')

# ---- PnP RANSAC --------------------------------------------------------------------------------------------------------
xyz = torch.randn(B, K, 3, device=dev)
xyz[..., 2] = xyz[..., 2].abs() + 2
kp2 = torch.stack([525 * xyz[..., 0] / xyz[..., 2] + 320 - 0.5, 525 * xyz[..., 1] / xyz[..., 2] + 240 - 0.5], -1).contiguous()
mt = torch.arange(K, device=dev).repeat(B, 1)
report('ransac_hyp_kernel + ransac_finalize_kernel (1024 hypotheses x 1024 points, fp64)',
       timeit(lambda: ops.ransac_pnp(kp2, mt, xyz, 525, 525, 320, 240, 8.0)), flops=B * 1024 * 1024 * 40.0, bound='cuda-core',
       note='40 FLOP per hypothesis-point (SURVEY 8d) + P3P + LM refinement')

print('\n| kernel | ms / launch | algorithmic | achieved | frac of measured peak |')
print('|---|---|---|---|---|')
for r in rows:
    alg = f"{r['algorithmic_MB']:.1f} MB" if r['bound'] in ('hbm', 'latency') and 'algorithmic_MB' in r else f"{r.get('algorithmic_GFLOP', 0):.1f} GFLOP"
    ach = f"{r['GBps']:.0f} GB/s" if r['bound'] in ('hbm', 'latency') and 'GBps' in r else f"{r.get('TFLOPs', 0):.1f} TFLOP/s"
    fr = f"{r['frac']:.2f} ({'HBM' if r['bound'] == 'hbm' else 'bf16 tensor'})" if 'frac' in r else '-'
    print(f"| {r['kernel']} | {r['ms']:.3f} | {alg} | {ach} | {fr} |")
if OUT:
    Path(OUT).write_text(json.dumps({'frames_per_launch': B, 'peaks': {'hbm_gbs': HBM, 'bf16_tflops': TC}, 'kernels': rows}, indent=1))
