"""ResBlock 1x1 convolutions (256 -> 256 @120x160, B frames) in isolation: conv1 (no residual) against conv3 (residual from the
split-bf16 planes, ReLU), CUDA events, L2 flushed between repetitions; checks conv3 against torch fp32.
    python tools/bench_resconv.py [B]"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pram_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device('cuda')
torch.manual_seed(0)
h, w, c = 120, 160, 256
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def timeit(fn, reps=9, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


xf = torch.randn(B, h, w, c, device=dev)
rf = torch.randn(B, h, w, c, device=dev)
x, r = ops.split_bf16(xf, True), ops.split_bf16(rf, True)
wf = torch.randn(1, c, c, device=dev) * 0.05
wt = ops.split_bf16(wf, True)
bias = torch.randn(c, device=dev)
out = ops.conv_tc(x, wt, bias, 1, 1, True, 3, res_bf=r)['bf']
ref = torch.relu((x.hi.float() + x.lo.float()) @ wf[0].t() + bias + (r.hi.float() + r.lo.float()))
err = ((out.hi.float() + out.lo.float()) - ref).abs().max().item() / ref.abs().max().item()
res = {'frames': B, 'conv3_relerr_vs_torch_fp32': err}
res['conv1_ms'] = timeit(lambda: ops.conv_tc(x, wt, bias, 1, 1, True, 3))
res['conv3_res_bf_ms'] = timeit(lambda: ops.conv_tc(x, wt, bias, 1, 1, True, 3, res_bf=r))
res['conv3_res_bf_f32out_ms'] = timeit(lambda: ops.conv_tc(x, wt, bias, 1, 1, True, 3, res_bf=r, want_f32=True))
res['conv3_res_f32_ms'] = timeit(lambda: ops.conv_tc(x, wt, bias, 1, 1, True, 3, res=rf))
by = B * h * w * c * 4
res['conv1_GBps'] = 2 * by / res['conv1_ms'] / 1e6
res['conv3_GBps'] = 3 * by / res['conv3_res_bf_ms'] / 1e6
print(json.dumps(res))
assert err < 2e-4, err
