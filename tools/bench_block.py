"""Fused transformer-block tail (csrc/mlp_block_tc.cu) vs the four separate launches it replaces, in isolation at the
bench token counts (T = 32 x 1024 for SegNetViT, 2 x 32 x 1024 for GML), CUDA events, L2 flushed between repetitions.

    python tools/bench_block.py [out.json]
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pram_b200 import ops  # noqa: E402
from pram_b200.nets import _blocks as B  # noqa: E402

dev = torch.device('cuda')
torch.manual_seed(0)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
peaks = json.loads((ROOT / 'MEASURED_PEAKS.json').read_text()) if (ROOT / 'MEASURED_PEAKS.json').exists() else {}
TC = float(peaks.get('bf16_tflops', 1590.0))


def timeit(fn, reps=9, warm=3):
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


blk = B.SelfBlockParams().to(dev)
pk = B.pack_self(blk)
rows = []
for T in (1024, 32768, 65536):
    for split in (3, 1):
        ws = B.Workspace(T, dev, split)
        ws.x.normal_()
        ops.split_bf16_into(ws.x.contiguous(), ws.x_bf) if False else None
        ws.ctx_in_bf = True
        for fused in (True, False):
            ws.fused = fused
            ms = timeit(lambda: B._finish_block(ws, pk))
            flops = 2.0 * T * (256 * 256 + 512 * 512 + 512 * 256)
            r = {'T': T, 'split': split, 'fused': fused, 'ms': ms, 'algorithmic_GFLOP': flops / 1e9,
                 'TFLOPs_algorithmic': flops / ms / 1e9, 'frac_of_bf16_peak': flops / ms / 1e9 / TC,
                 'issued_TFLOPs': (3 if split == 3 else 1) * 2.0 * T * (512 * 512 + 512 * 256) / ms / 1e9 if fused else None}
            rows.append(r)
            print(json.dumps(r), flush=True)
if len(sys.argv) > 1:
    Path(sys.argv[1]).write_text(json.dumps(rows, indent=1))

# ---- timeline of one launch (SM clock stamps written by the kernel itself) -------------------------------------
if '--timeline' in sys.argv:
    T, split = 32768, 3
    ws = B.Workspace(T, dev, split)
    ws.ctx_in_bf = True
    dbg = torch.zeros(148 * 8 * 32, device=dev, dtype=torch.int64)
    cbf, nbf = ws.cat_bf[0], ws.cat_bf[1]
    for _ in range(3):
        ops.mlp_block_tc(cbf, 512, T, pk['blk.w1.tc'], pk['blk.w3.tc'], pk['blk.tables'], None, 0, None, 0, nbf, 512, split=split, dbg=dbg)
    torch.cuda.synchronize()
    d = dbg.view(148, 8, 32).cpu()
    names = ['epi: wait h_full', 'epi: h_full', 'epi: stats done', 'epi: pass2 done', 'epi: acc_full', 'epi: phaseC done', '', '',
             'mma: first stage landed', 'mma: gemm1 issued', 'mma: gemm2 issued']
    for cta in (0, 73, 147):
        t0 = int(d[cta, 0, 0])
        for it in range(3):
            if int(d[cta, it, 0]) == 0:
                continue
            print(f'cta {cta} tile {it}: ' + ', '.join(f'{names[k]}={int(d[cta, it, k]) - t0}' for k in (0, 8, 9, 1, 2, 3, 10, 4, 5)))
            print('    epi k-block written :', [int(d[cta, it, 16 + j]) - t0 for j in range(8)])
            print('    mma k-block issued  :', [int(d[cta, it, 24 + j]) - t0 for j in range(8)])
