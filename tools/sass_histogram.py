"""Opcode histogram per .cu of the shipped library (cuobjdump -sass on the per-file objects): the evidence the review
asks for -- UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMAPF = TMA load / prefetch, HMMA = legacy
mma.sync.  usage: python tools/sass_histogram.py > profiles/r02_sass_opcodes.txt   (CPU only; needs the built objects)"""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
KEY = ('UTCHMMA', 'UTCQMMA', 'UTCBAR', 'UTCATOM', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'UBLKCP', 'HMMA', 'LDSM', 'LDGSTS',
       'SYNCS', 'ELECT', 'FENCE', 'UCGABAR', 'MUFU', 'ATOMG', 'ATOMS', 'RED')
print(f'# cuobjdump -sass opcode histogram per source file (sm_100a), objects under pram_b200/csrc/build/')
tot = Counter()
for obj in sorted((ROOT / 'pram_b200' / 'csrc' / 'build').glob('*.o')):
    sass = subprocess.run(['cuobjdump', '-sass', str(obj)], capture_output=True, text=True).stdout
    ops = Counter()
    kernels = 0
    for line in sass.splitlines():
        if 'Function :' in line:
            kernels += 1
        m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m:
            ops[m.group(1)] += 1
    n = sum(ops.values())
    key = {k: sum(v for o, v in ops.items() if o.startswith(k)) for k in KEY}
    tot.update(key)
    print(f'\n## {obj.stem}.cu   ({kernels} kernels, {n} SASS instructions)')
    print('   key ops : ' + ', '.join(f'{k} {v}' for k, v in key.items() if v))
    print('   top     : ' + ', '.join(f'{o} {v}' for o, v in ops.most_common(14)))
print('\n## whole library, key opcodes\n   ' + ', '.join(f'{k} {v}' for k, v in tot.items() if v))
