"""bench.py contract on CPU: the reference arm prints ONE JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

from conftest import REPO


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS='4')
    r = subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), '--impl', 'reference', '--workload', 'cfg1', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600, env=env, cwd=REPO)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'molecules/sec (T=1000)' and d['unit'] == 'molecules/s'
    assert d['higher_is_better'] is True and d['value'] > 0 and d['steps'] == 1
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'molecules/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=120, env=env, cwd=REPO)
    assert r.returncode == 0 and r.stdout.strip() == ''


import pytest


@pytest.mark.gpu
def test_bench_line_on_gpu():
    """The product arm on cuda:0 (short run): one JSON line with the contract's keys, roofline and per-kernel table."""
    r = subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), '--steps', '3', '--warmup', '3', '--e2e-steps', '16', '--no-cpu-baseline'],
                       capture_output=True, text=True, timeout=600, cwd=REPO)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype',
                'data', 'config', 'e2e', 'gpu_launches', 'clocks', 'roofline'):
        assert key in d, key
    assert d['n_gpus'] == 1 and d['steps'] == 3 and d['warmup'] >= 3 and d['value'] > 0 and d['gpu_launches'] > 0
    assert d['roofline']['bound'] == 'hbm' and 0 < d['roofline']['frac'] < 1.5 and d['roofline']['unit'] == 'GB/s'
    assert d['e2e']['value'] > 0 and d['e2e']['h2d_bytes_per_step'] > 0 and d['e2e']['d2h_bytes_per_step'] > 0
    assert 'workload' in d['config'] and 'model' not in d['config']
