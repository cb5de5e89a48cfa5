"""ResBlock 1x1 convs and the L2-normalised descriptor projection in isolation, for ncu / A-B timing:
   conv4.x.conv1 (256->256), conv4.x.conv3 (256->256 + residual planes), convDb (256->128, fp16 plane in, fp32 + L2 norm out).
   python tools/profile_trunk_tail.py [B] [time]"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pram_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device('cuda')
torch.manual_seed(0)
h, w, c = 120, 160, 256
xf = torch.randn(B, h, w, c, device=dev)
x, r = ops.split_bf16(xf, True), ops.split_bf16(torch.randn(B, h, w, c, device=dev), True)
wt = ops.split_bf16(torch.randn(1, c, c, device=dev) * 0.05, True)
bias = torch.randn(c, device=dev)
x16 = ops.as_f16_plane(xf)
wd = ops.Split(((torch.randn(1, 128, c, device=dev) * 0.05).half()).view(torch.bfloat16), None)
bd = torch.randn(128, device=dev)


def c1(): return ops.conv_tc(x, wt, bias, 1, 1, True, 3)
def c3(): return ops.conv_tc(x, wt, bias, 1, 1, True, 3, res_bf=r)
def c3last(): return ops.conv_tc(x, wt, bias, 1, 1, True, 3, res_bf=r, want_f32=True, want_ps=True)
def db(): return ops.conv_tc(x16, wd, bd, 1, 1, False, 1, want_f32=True, want_bf=False, l2norm=True, f16=True)
def db_nonorm(): return ops.conv_tc(x16, wd, bd, 1, 1, False, 1, want_f32=True, want_bf=False, l2norm=False, f16=True)


cases = [('c1', c1), ('c3', c3), ('c3last', c3last), ('db', db), ('db_nonorm', db_nonorm)]
if len(sys.argv) > 2:
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
    res = {}
    for cl in (1, 2):
        ops.GEMM_CLUSTER = cl
        for name, fn in cases:
            res[f'{name}_cl{cl}_ms'] = round(timeit(fn), 4)
    ops.GEMM_CLUSTER = 0
    ops.GEMM_L2_PREFETCH = 1
    for name, fn in cases[:2]:
        res[f'{name}_l2pf_ms'] = round(timeit(fn), 4)
    print(json.dumps(res))
else:
    for _ in range(2):
        for name, fn in cases:
            fn()
    torch.cuda.synchronize()
    print('done')
