"""Per-tile cost of gemm_tc_kernel on Linear shapes: time vs number of tile rounds (rows = rounds x 148 x 128 / n_tiles),
for different epilogue outputs.  Slope = steady-state time per round, intercept = fixed launch/fill/drain cost."""
import sys, torch
sys.path.insert(0, '.')
from pram_b200 import ops
dev = torch.device('cuda')
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def run(rows, k, n, outs, reps=5):
    a = ops.split_bf16(torch.randn(rows, k, device=dev), True)
    w = ops.split_bf16(torch.randn(n, k, device=dev) * 0.05, True)
    bias = torch.randn(n, device=dev)
    of = torch.empty(rows, n, device=dev) if 'f' in outs else None
    ob = ops.empty_split((rows, n), dev, True) if 'b' in outs else None
    res = torch.randn(rows, n, device=dev) if 'r' in outs else None
    fn = lambda: ops.linear_tc(a, k, rows, k, w, n, bias, res, n if res is not None else 0, False, of, n, ob, n, split=3)
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]


for k, n in ((256, 256), (512, 512), (512, 256)):
    nt = (n + 255) // 256
    for outs in ('', 'f', 'b', 'fb', 'fbr'):
        line = []
        for rounds in (1, 2, 4, 8):
            rows = rounds * 148 * 128 // nt
            line.append(f'{run(rows, k, n, outs):7.1f}')
        print(f'K={k} N={n} outs={outs or "-":4s} us for 1/2/4/8 rounds: ' + ' '.join(line), flush=True)
