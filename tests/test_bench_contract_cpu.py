"""CPU: the reference arm of bench.py (the oracle port on the host cores)
prints exactly one JSON line with the keys the driver reads; the product arm
refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'),
                           *args], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    res = _run('--impl', 'reference', '--steps', '1', '--warmup', '0')
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference'
    assert d['metric'] == 'dmc_walker_steps_per_sec'
    assert d['unit'] == 'walker-steps/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['n_gpus'] == 1 and d['steps'] == 1
    assert d['dtype'] == 'f64' and d['data'] == 'synthetic'
    assert 'workload' in d['config'] and 'model' not in d['config']
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['sample']
    assert cb['value'] == d['value']
    e2e = d['e2e']
    assert e2e['value'] == d['value'] and e2e['unit'] == d['unit']
    assert e2e['h2d_bytes_per_step'] == 0 and e2e['d2h_bytes_per_step'] == 0


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip('a CUDA device is present')
    res = _run('--steps', '1', '--warmup', '1', timeout=300)
    assert res.returncode != 0
    assert 'no CUDA device' in (res.stderr + res.stdout)
    assert not [ln for ln in res.stdout.splitlines() if ln.startswith('{')]
