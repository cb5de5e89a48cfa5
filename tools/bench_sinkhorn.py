"""Sinkhorn + mutual matches alone: B pairs of 1024 x 1024, 20 iterations, CUDA events, L2 flushed between repetitions.
    python tools/bench_sinkhorn.py [B]            (PRAM_SINKHORN_MAXG=8 keeps the portable cluster size)"""
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pram_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device('cuda')
torch.manual_seed(0)
K = 1024
dist = torch.randn(B, K, K, device=dev) * 3
bin_score = torch.tensor(1.0, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
ts = []
for r in range(12):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = ops.sinkhorn_match(dist, bin_score, 20, 0.2); e1.record()
    torch.cuda.synchronize()
    if r >= 3:
        ts.append(e0.elapsed_time(e1))
ts.sort()
print(json.dumps({'B': B, 'maxg': os.environ.get('PRAM_SINKHORN_MAXG', 'auto'), 'ms': round(ts[len(ts) // 2], 4),
                  'checksum': int(out[0].sum().item())}))
