#!/usr/bin/env python
"""profiles/rNN_ncu_top_kernel.json from an `ncu --set full` capture of one forward (raw-page CSV or .ncu-rep):
per kernel class (as bench.py names them) the launches, average duration, DRAM bytes per launch and tensor-pipe
utilisation, plus the per-launch rows.  bench.py reads `dram_bytes_per_launch` of the dominant class as
`roofline.traffic`.

    python tools/ncu_top_kernel.py profiles/r02_ncu_top_kernel.json gpurun_out/r02_ncu_full_fwd32_raw.csv "source note"
"""
import json
import sys

from ncu_summary import rows_of

CLASS = (('k_tcn', 'gemm_tcn'), ('k_gcnw', 'gemm_1x1'), ('k_gcn_tc2', 'gemm_1x1'), ('k_ln_', 'frame'), ('k_rt_', 'frame'))


def main():
    out, src = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ''
    rows = rows_of(src)
    res = {}
    for r in rows:
        cls = next((c for k, c in CLASS if k in r['kernel']), None)
        if cls is None:
            continue
        d = res.setdefault(cls, {'dram_bytes_per_launch': 0.0, 'launches': 0, 'avg_duration_us': 0.0,
                                 'avg_tensor_pipe_pct': 0.0})
        d['launches'] += 1
        d['dram_bytes_per_launch'] += r.get('dram_bytes_per_launch', 0.0)
        d['avg_duration_us'] += r.get('duration_us') or 0.0
        d['avg_tensor_pipe_pct'] += r.get('tensor_pipe_pct', 0.0)
    for d in res.values():
        for k in ('dram_bytes_per_launch', 'avg_duration_us', 'avg_tensor_pipe_pct'):
            d[k] /= d['launches']
    res['source'] = note
    res['per_launch'] = rows
    json.dump(res, open(out, 'w'), indent=1)
    for k, d in res.items():
        if isinstance(d, dict):
            print(k, d)


if __name__ == '__main__':
    main()
