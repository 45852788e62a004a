#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (tcgen05 MMA = UTC*MMA, TMEM
load/store = LDTM/STTM, TMA = UTMALDG/UTMASTG/UBLKCP, legacy warp MMA = HMMA, cp.async = LDGSTS), from
`cuobjdump -sass` of the built library (no GPU needed).

    python tools/sass_summary.py [lib.so] > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'realtime-st-gcn_b200', 'csrc', 'libstgcn_b200.so')
WANT = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTCBAR', 'SYNCS', 'HMMA', 'LDGSTS',
        'CCTL.IVALL', 'MEMBAR', 'ATOMG', 'REDG']
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()
kern, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m:
        op = m.group(1)
        total[kern] += 1
        for w in WANT:
            if op.startswith(w):
                counts[kern][w] += 1
print("# SASS mnemonic counts per kernel of %s (cuobjdump -sass; sm_100a)" % os.path.basename(lib))
print("# UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UBLKCP = TMA tensor / bulk copies, HMMA = mma.sync")
for k, c in counts.items():
    if not total[k]:
        continue
    name = demangle(k)
    name = re.sub(r'\(.*', '', name).replace('stgcn::', '').replace('void ', '')
    tags = '  '.join('%s %d' % (w, c[w]) for w in WANT if c[w])
    print('%-52s %6d instr  %s' % (name[:52], total[k], tags))
