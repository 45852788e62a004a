#!/usr/bin/env python
"""Summarise `ncu --set full` reports for profiles/: per kernel the duration, DRAM bytes, tensor-pipe
and issue utilisation, registers and shared memory.

    python tools/ncu_summary.py out.json label=report.ncu-rep [label=report2.ncu-rep ...]
Writes a JSON {label: [ {kernel, duration_us, dram_bytes_per_launch, ...}, ... ]} and prints a table."""
import csv
import io
import json
import subprocess
import sys

WANT = {
    'gpu__time_duration.sum': 'duration',
    'dram__bytes_read.sum': 'dram_read',
    'dram__bytes_write.sum': 'dram_write',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_pipe_pct',
    'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_active_pct',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed': 'l2_pct',
    'launch__registers_per_thread': 'registers',
    'launch__shared_mem_per_block_dynamic': 'dyn_smem',
    'launch__grid_size': 'grid',
    'launch__block_size': 'block',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct',
}
SCALE = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}


def rows_of(rep):
    # a .ncu-rep, or the `ncu -i rep --page raw --csv` dump of one (written on the GPU box when the report itself is too large to bring back)
    if rep.endswith('.csv'):
        out = open(rep).read()
    else:
        out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {'kernel': r[hdr.index('Kernel Name')][:60]}
        for m, k in WANT.items():
            if m in hdr:
                i = hdr.index(m)
                try:
                    v = float(r[i].replace(',', ''))
                except ValueError:
                    continue
                d[k] = v * SCALE.get(units[i], 1)
        d['duration_us'] = d.pop('duration', None)
        if 'dram_read' in d:
            d['dram_bytes_per_launch'] = d['dram_read'] + d['dram_write']
        res.append(d)
    return res


def main():
    out_path, result = sys.argv[1], {}
    for arg in sys.argv[2:]:
        label, rep = arg.split('=', 1)
        result[label] = rows_of(rep)
        for d in result[label]:
            print("%-10s %-58s %9.1f us  dram %8.1f MB  tensor %5.1f%%  issue %5.1f%%  dram %5.1f%%  L2 %5.1f%%  regs %d" % (
                label, d['kernel'], d.get('duration_us') or 0, d.get('dram_bytes_per_launch', 0) / 1e6,
                d.get('tensor_pipe_pct', 0), d.get('issue_active_pct', 0), d.get('dram_pct', 0), d.get('l2_pct', 0),
                int(d.get('registers', 0))))
    json.dump(result, open(out_path, 'w'), indent=1)


if __name__ == '__main__':
    main()
