"""Quick on-GPU probe: parity of model_eval against the oracle, fp64 peak and
DMC step-kernel throughput at a few sizes.  Development aid, not the bench."""
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, '.')
sys.path.insert(0, 'oracle')
from phd_qmclib_b200 import engine, model  # noqa: E402
import oracle  # noqa: E402

PI = math.pi
print('fp64 burst peak %.2f TF' % engine.measure_fp64_peak(0))

# parity spot check (model_eval) at N=100 and an awkward N
for nop, L, rm in [(100, 100.0, 25.0), (21, 17.5, 1.75), (7, 10.0, 3.3)]:
    spec = model.Spec(5 * PI ** 2, 1, 2, nop, L, rm)
    p = model.param_block(spec)
    rng = np.random.default_rng(5)
    confs = np.zeros((512, 2, nop))
    confs[:, 0] = rng.random((512, nop)) * L
    with engine.Engine(spec) as eng:
        o = eng.model_eval(confs)
    r = oracle.model_eval(p, confs)
    sc = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), np.median(np.abs(b)))))
    print(f'N={nop}: lnpsi {sc(o["lnpsi"], r["lnpsi"]):.2e} energy '
          f'{sc(o["energy"], r["energy"]):.2e} drift '
          f'{np.max(np.abs(o["drift"] - r["drift"])) / np.max(np.abs(r["drift"])):.2e}')

cases = [(100, 125000, False), (50, 100000, False), (200, 20000, False),
         (20, 100000, False)]
if len(sys.argv) > 1:
    cases = [(100, 125000, False), (100, 125000, True)]
for nop, nw, shuffled in cases:
    spec = model.Spec(5 * PI ** 2, 1, 2, nop, nop, 0.25 * nop)
    eng = engine.Engine(spec)
    rng = np.random.default_rng(0)
    ini = np.zeros((nw, 2, nop))
    ini[:, 0] = np.arange(nop)[None, :] + 0.25 + 0.15 * (rng.random((nw, nop)) - 0.5)
    if shuffled:        # same physics, particle labels in random order
        ini[:, 0] = rng.permuted(ini[:, 0], axis=1)
    cap = int(nw * 1.25)
    dp = eng.dmc_params(6.25e-4, cap, nw, 0.5, 7, 0.0, float(nop))
    eng.dmc_init(dp, ini)
    eng.dmc_run_block(16)
    eng.set_profiling(True)
    nts = 32
    out = eng.dmc_run_block(nts)
    st = eng.last_block_stats()
    ws = float(out['num_walkers'].sum())
    F = 58 * nop * (nop - 1) / 2 + 113 * nop + 35
    print(f'N={nop} W={nw} shuffled={shuffled} KC={os.environ.get("QMCB_KC")} NT={os.environ.get("QMCB_NT")}: '
          f'block {st["total_ms"]:.2f} ms step-kernel {st["step_kernel_ms"]:.2f} ms  '
          f'{ws / (st["total_ms"] * 1e-3):.3e} ws/s  '
          f'alg {ws * F / (st["step_kernel_ms"] * 1e-3) / 1e12:.2f} TF  '
          f'nw[-1]={int(out["num_walkers"][-1])} E/N='
          f'{out["energy"][-1] / out["weight"][-1] / nop:.4f}', flush=True)
    eng.close()
