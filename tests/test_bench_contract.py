"""bench.py's reference arm (the oracle port timed on the host cores) runs without a GPU: its JSON line must carry
the keys the driver reads (metric / unit / config shared with the B200 arm, impl, cpu_baseline, e2e)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '1', '--frames', '120', '--trials', '2'], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference'
    assert line['metric'] == 'stgcn_fwd_skeleton_frames_per_s' and line['unit'] == 'frames/s'
    assert line['higher_is_better'] is True and line['value'] > 0 and line['steps'] == 1
    assert line['config']['frames_per_trial'] == 120 and 'workload' in line['config']
    cb = line['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == line['value'] and cb['sample']
    assert line['e2e'] == {'value': line['value'], 'unit': line['unit'], 'h2d_bytes_per_step': 0,
                           'd2h_bytes_per_step': 0}


def test_b200_arm_refuses_measurement_builds():
    """STGCN_DEBUG / STGCN_LIB select the measurement build: bench.py must not produce a bench line with them."""
    env = dict(os.environ, STGCN_DEBUG='4')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1', '--warmup', '1'], env=env,
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0
    assert 'STGCN_DEBUG' in (out.stderr + out.stdout)
