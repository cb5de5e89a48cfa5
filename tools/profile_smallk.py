"""Small-K GEMM (ResBlock 1x1 conv shape: M = B*120*160, K = 256, N = 256, fp32 out) for ncu --set full."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pram_b200 import ops
dev = torch.device('cuda')
B = 32
x = ops.split_bf16(torch.randn(B, 120, 160, 256, device=dev))
w = ops.split_bf16(torch.randn(1, 256, 256, device=dev) * 0.05)
bias = torch.randn(256, device=dev)
for _ in range(3):
    ops.conv_tc(x, w, bias, 1, 1, True, 3, want_f32=True, want_bf=False)
torch.cuda.synchronize()
print('done')
