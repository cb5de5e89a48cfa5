"""Fused-qkv GEMM (T=32768 tokens, K=256, N=768 -> split-bf16 Q/K/V) in isolation for ncu --set full."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pram_b200 import ops
dev = torch.device('cuda')
T, B, N = 32768, 32, 1024
x = ops.split_bf16(torch.randn(T, 512, device=dev))
w = ops.split_bf16(torch.randn(768, 256, device=dev) * 0.05)
bias = torch.randn(768, device=dev)
cos = torch.rand(T, 32, device=dev); sin = torch.rand(T, 32, device=dev)
q, k, v = (ops.empty_split((T, 256), dev) for _ in range(3))
qkv = {'mode': 1, 'scale': 1.0, 'cos': cos, 'sin': sin, 'q': q, 'k': k, 'v': v, 'seg_split': T, 'seg_n0': N, 'seg_n1': N}
for _ in range(3):
    ops.linear_tc(x, 512, T, 256, w, 768, bias, split=3, bn=256, qkv=qkv)
out = torch.empty(T, 768, device=dev)
for _ in range(3):
    ops.linear_tc(x, 512, T, 256, w, 768, bias, out_f32=out, ld_f32=768, split=3, bn=256)
torch.cuda.synchronize()
print('done')
