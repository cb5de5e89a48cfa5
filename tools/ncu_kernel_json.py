"""One kernel of an `ncu --set full` report as a small JSON (the numbers bench.py's roofline block and profiles/README.md quote).
usage: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_kernel_json.py <kernel-substring> <launch description> > out.json"""
import csv
import json
import sys

rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
want = sys.argv[1]
r = next(r for r in rows[2:] if want in r[hdr.index('Kernel Name')])
g = lambda k: float(r[hdr.index(k)].replace(',', ''))
u = lambda k: units[hdr.index(k)]
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
rd = g('dram__bytes_read.sum') * scale[u('dram__bytes_read.sum')]
wr = g('dram__bytes_write.sum') * scale[u('dram__bytes_write.sum')]
tscale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}
out = {'kernel': r[hdr.index('Kernel Name')], 'launch': sys.argv[2] if len(sys.argv) > 2 else '',
       'gpu_time_ms': g('gpu__time_duration.sum') * tscale[u('gpu__time_duration.sum')],
       'dram_bytes_read': rd, 'dram_bytes_write': wr, 'dram_bytes': rd + wr,
       'tensor_pipe_active_pct': g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
       'registers': g('launch__registers_per_thread'),
       'l2_to_sm_bytes': g('lts__t_bytes.sum') * scale.get(u('lts__t_bytes.sum'), 1.0) if 'lts__t_bytes.sum' in hdr else None,
       'source': 'ncu --set full --clock-control none --import-source on (gpurun_out/*.ncu-rep, not tracked)'}
print(json.dumps(out, indent=1))
