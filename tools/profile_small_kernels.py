"""The non-GEMM kernels of the hot path, one warm-up + one launch each at the bench shape (32 frames of 640x480, K = 1024), for
`ncu --set full` (profiles/README.md): conv1a, grouped conv, score map, NMS, selection, sampling, positional encoding,
segmentation ranking, Sinkhorn, RANSAC, the fused block tail and the qkv GEMM.
    ncu --set full --clock-control none -o out python tools/profile_small_kernels.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pram_b200 import ops  # noqa: E402
from pram_b200.nets import _blocks as B  # noqa: E402

Bn, H, W, K = 32, 480, 640, 1024
dev = torch.device('cuda')
torch.manual_seed(0)
img = torch.rand(Bn, 3, H, W, device=dev)
w1, b1 = torch.randn(27, 64, device=dev) * 0.1, torch.randn(64, device=dev) * 0.1
xg = ops.split_bf16(torch.randn(Bn, 120, 160, 256, device=dev), True)
wg, bg = torch.randn(9 * 8 * 8 * 32, device=dev) * 0.1, torch.randn(256, device=dev)
logits = torch.randn(Bn, H // 8, W // 8, 65, device=dev) * 3
dmap = torch.randn(Bn, 120, 160, 128, device=dev)
lg = torch.randn(Bn, K, 113, device=dev)
dist = torch.randn(Bn, K, K, device=dev) * 3
bin_score = torch.tensor(1.0, device=dev)
xyz = torch.randn(Bn, K, 3, device=dev)
xyz[..., 2] = xyz[..., 2].abs() + 2
kp2 = torch.stack([525 * xyz[..., 0] / xyz[..., 2] + 320 - 0.5, 525 * xyz[..., 1] / xyz[..., 2] + 240 - 0.5], -1).contiguous()
mt = torch.arange(K, device=dev).repeat(Bn, 1)
wr = torch.randn(32, 2, device=dev)
blk = B.SelfBlockParams().to(dev)
pk = B.pack_self(blk)
ws = B.Workspace(Bn * K, dev, 3)
ws.ctx_in_bf = True
for rep in range(2):
    ops.conv1a(img, w1, b1, 3)
    ops.gconv3x3_tc(xg, wg, bg, True, 3)
    score = ops.score_map(logits)
    kp, sc, n, _ = ops.detect_keypoints(score, 0.005, 128, K, 4)
    ops.sample_features(dmap, kp, n, 4, True)
    ops.posenc(kp, W, H, wr)
    ops.rank_landmarks(lg, None, 20, 8)
    ops.sinkhorn_match(dist, bin_score, 20, 0.2)
    ops.ransac_pnp(kp2, mt, xyz, 525, 525, 320, 240, 8.0)
    B._finish_block(ws, pk)
torch.cuda.synchronize()
print('done')
