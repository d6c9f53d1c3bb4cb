"""CPU tier: the bench.py contract that can be checked without a GPU -- the reference arm
(`--impl reference`) prints exactly ONE JSON line on stdout with the keys the driver reads, and
non-zero ranks of a multi-rank launch exit without work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env, *args):
    env = dict(os.environ, B2N_BENCH_CPU_SECONDS='0.3', OMP_NUM_THREADS='1', **extra_env)
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), *args], cwd=ROOT, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


def test_reference_arm_json_line():
    out = _run({}, '--impl', 'reference', '--gpus', '1', '--steps', '1', '--warmup', '0')
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'proposals/s' and d['higher_is_better'] is True
    assert d['n_gpus'] == 1 and d['steps'] == 1 and d['value'] > 0 and d['gpu_launches'] == 0
    cb = d['cpu_baseline']
    assert cb['kind'] in ('reference', 'port') and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
    e = d['e2e']
    assert e['value'] == d['value'] and e['h2d_bytes_per_step'] == 0 and e['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config'] and 'model' not in d['config']
    assert d['metric'].startswith('rwalk proposals/sec')


def test_reference_arm_other_ranks_are_silent():
    out = _run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'}, '--impl', 'reference', '--gpus', '2', '--steps', '1',
               '--warmup', '0')
    assert out.strip() == ''
