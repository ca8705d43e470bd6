"""CPU: the reference arm of bench.py (the live Numba reference on the host
cores; the oracle port, with the reason, when no reference tree is found)
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


def _check_line(res, metric, unit):
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference'
    assert d['metric'] == metric
    assert d['unit'] == unit and d['higher_is_better'] is True
    assert d['value'] > 0 and d['n_gpus'] == 1 and d['steps'] == 1
    assert d['dtype'] == 'f64' and d['data'] == 'synthetic'
    assert 'workload' in d['config'] and 'model' not in d['config']
    cb = d['cpu_baseline']
    assert cb['cores'] >= 1 and cb['sample']
    assert cb['value'] == d['value']
    e2e = d['e2e']
    assert e2e['value'] == d['value'] and e2e['unit'] == d['unit']
    assert e2e['h2d_bytes_per_step'] == 0 and e2e['d2h_bytes_per_step'] == 0
    return d


def test_reference_arm_falls_back_to_the_port_with_a_reason(tmp_path):
    """No reference tree to be found: the C port of oracle/ is timed and the
    line says why."""
    env = dict(os.environ, QMCB_REFERENCE_SRC=str(tmp_path),
               QMCB_REFERENCE_ONLY_ENV='1')
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'),
                          '--impl', 'reference', '--steps', '1', '--warmup',
                          '0'], capture_output=True, text=True, timeout=600,
                         cwd=ROOT, env=env)
    d = _check_line(res, 'dmc_walker_steps_per_sec', 'walker-steps/s')
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port'
    assert 'reference' in cb['reason'] and 'unavailable' in cb['reason']


def test_reference_arm_times_the_live_reference():
    """With a reference tree (builder container: /root/reference; GPU box:
    baseline/_ref) the arm runs mrbp_qmc.vmc.Sampling.blocks() under Numba
    (the VMC config: its JIT compilation is the quick one)."""
    import pytest
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_arm
    ok, why = ref_arm.probe()
    if not ok:
        pytest.skip(why)
    res = _run('--impl', 'reference', '--config', 'c2_vmc', '--steps', '1',
               '--warmup', '0')
    d = _check_line(res, 'vmc_chain_steps_per_sec', 'chain-steps/s')
    cb = d['cpu_baseline']
    assert cb['kind'] == 'reference' and cb['cores'] == 1
    assert 'reason' not in cb and cb['numba']
    assert 'mrbp_qmc.vmc.Sampling' in cb['sample']


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip('a CUDA device is present')
    res = _run('--steps', '1', '--warmup', '1', timeout=300)
    assert res.returncode != 0
    assert 'no CUDA device' in (res.stderr + res.stdout)
    assert not [ln for ln in res.stdout.splitlines() if ln.startswith('{')]
