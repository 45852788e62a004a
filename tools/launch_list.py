#!/usr/bin/env python
"""Print an `ncu --csv --metrics ...` launch list as one line per launch (development aid)."""
import csv, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
iK, iM, iV, iID = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
d = {}
for r in rows[1:]:
    d.setdefault(r[iID], {'k': r[iK][:44]})[r[iM].split('.')[0]] = r[iV]
tot = 0.0
for i, (k, v) in enumerate(d.items()):
    us = float(v.get('gpu__time_duration', '0').replace(',', '')) / 1e3
    tot += us
    mb = lambda key: ('%8.1f' % (float(v[key].replace(',', '')) / 1e6)) if key in v else '       -'
    print("%3d %-44s %9.1f us  tensor %5s%%  lts %5s%%  dram %5s%%  issue %5s%%  dram rd/wr MB %s %s" % (
        i, v['k'], us, v.get('sm__pipe_tensor_cycles_active', '-'), v.get('lts__throughput', '-'),
        v.get('dram__throughput', '-'), v.get('smsp__issue_active', '-'), mb('dram__bytes_read'), mb('dram__bytes_write')))
print("total %.1f us" % tot)
