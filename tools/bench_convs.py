"""Every tcgen05 convolution of the SFD2 trunk at the bench shape (B frames of 640x480), bf16x3, alone: CUDA events, L2 flushed
between repetitions; 2-CTA clusters with a multicast weight tile (PRAM_GEMM_CLUSTER=2) against single CTAs.
    python tools/bench_convs.py [B]"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pram_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device('cuda')
torch.manual_seed(0)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def timeit(fn, reps=7, warm=2):
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


def case(name, h, w, cin, cout, stride, bn=0):
    """h, w = INPUT size; stride 2 reads the 2x2 phase-split planes like the trunk does."""
    wt = ops.split_bf16(torch.randn(9, cout, cin, device=dev) * 0.02, True)
    bias = torch.randn(cout, device=dev)
    if stride == 1:
        x = ops.split_bf16(torch.randn(B, h, w, cin, device=dev), True)
        fn = lambda: ops.conv_tc(x, wt, bias, 3, 1, True, 3, bn=bn)
        ho, wo = h, w
    else:
        ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        x = ops.split_bf16(torch.randn(B * 4, ho, wo, cin, device=dev), True)
        fn = lambda: ops.conv_tc(x, wt, bias, 3, 2, True, 3, out_shape_hw=(ho, wo), bn=bn)
    r = {'conv': name}
    for cl in (1, 2):
        ops.GEMM_CLUSTER = cl
        r[f'cl{cl}_ms'] = round(timeit(fn), 4)
    ops.GEMM_CLUSTER = 0
    fl = 2.0 * B * ho * wo * cout * cin * 9
    r['issue_bound_ms'] = round(3 * fl / 1694.7e12 * 1e3, 4)
    print(json.dumps(r), flush=True)


case('conv1b 64->64 s2 @480x640', 480, 640, 64, 64, 2)
case('conv2a 64->128 @240x320', 240, 320, 64, 128, 1)
case('conv2b 128->128 s2 @240x320', 240, 320, 128, 128, 2)
case('conv3a 128->256 @120x160', 120, 160, 128, 256, 1)
case('conv3b 256->256 @120x160', 120, 160, 256, 256, 1)
case('convPa.0 256->256 s2 @120x160', 120, 160, 256, 256, 2)
case('conv2a bn=64', 240, 320, 64, 128, 1, bn=64)
