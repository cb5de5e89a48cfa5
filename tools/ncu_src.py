"""Summarise the source page of an .ncu-rep: opcode mix, top stall sites, headline metrics.
usage: python tools/ncu_src.py <rep> [n_top]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__cycles_active.avg', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'launch__grid_size', 'launch__block_size', 'sm__maximum_warps_per_active_cycle_pct']
for h, u, v in zip(hdr, units, vals):
    if h in want or h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued'):
        print(f'{h:80s} {u:12s} {v}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); tot = 0; top = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    s = r[ix['Source']].strip(); parts = s.split()
    op = parts[1] if parts and parts[0].startswith('@') and len(parts) > 1 else (parts[0] if parts else '?')
    n = int(r[ix['Instructions Executed']]); ops[op.split('.')[0]] += n; tot += n
    top.append((int(r[ix['# Samples']]), s[:70], n, {k: r[ix[k]] for k in ('stall_long_sb', 'stall_short_sb', 'stall_mio', 'stall_wait', 'stall_math', 'stall_barrier', 'stall_lg') if r[ix[k]] != '0'}))
print('\ntotal warp-instr', tot)
for op, n in ops.most_common(18): print(f'  {op:14s} {n:10d} {100*n/tot:5.1f}%')
top.sort(key=lambda t: -t[0])
print()
for t in top[:ntop]: print(t)
