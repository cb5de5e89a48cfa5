"""Launch the dominant kernels in isolation for `ncu --set full` (run under gpurun):
   conv3b-shaped tcgen05 implicit GEMM (B x 120 x 160, 256 -> 256, 3x3) and flash attention (B*4 heads, N=1024)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pram_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
split = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device('cuda')
torch.manual_seed(0)
x = ops.split_bf16(torch.randn(B, 120, 160, 256, device=dev), split == 3)
w = ops.split_bf16(torch.randn(9, 256, 256, device=dev) * 0.02, split == 3)
bias = torch.randn(256, device=dev)
for _ in range(3):
    ops.conv_tc(x, w, bias, 3, 1, True, split)
N = 1024
q = ops.split_bf16(torch.randn(B * 4, N, 64, device=dev), split == 3)
k = ops.split_bf16(torch.randn(B * 4, N, 64, device=dev), split == 3)
vt = ops.split_bf16(torch.randn(B * 4, 64, N, device=dev), split == 3)
out = ops.empty_split((B, N, 256), dev, split == 3)
for _ in range(3):
    ops.attention_tc(q, k, vt, B, 4, N, N, N, 0.125, None, out, 256, split)
torch.cuda.synchronize()
print('done')
