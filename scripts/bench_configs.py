"""Throughput of the other BASELINE.json configs on one GPU (development aid;
the headline line is bench.py).  Prints one JSON object per config."""
import json
import math
import sys
import time

import numpy as np

sys.path.insert(0, '.')
from phd_qmclib_b200 import engine, model  # noqa: E402

PI = math.pi


def lattice_ini(nw, nop, seed):
    rng = np.random.default_rng(seed)
    ini = np.zeros((nw, 2, nop))
    ini[:, 0] = np.arange(nop)[None, :] + 0.25 + 0.15 * (rng.random((nw, nop)) - 0.5)
    return ini


def vmc_c2(nch=100000, nop=50, ns=256, nblocks=3):
    spec = model.Spec(5 * PI ** 2, 1, 4, nop, nop, 0.25 * nop)
    eng = engine.Engine(spec)
    eng.vmc_init(lattice_ini(nch, nop, 0), 0.25 * spec.well_width, 1, 0.0, float(nop),
                 ssf_num_modes=nop)
    out = []
    for b in range(nblocks):
        t0 = time.perf_counter()
        o = eng.vmc_run_block(ns, series=False, sums=True)
        wall = time.perf_counter() - t0
        ms = eng.last_block_stats()['total_ms']
        out.append(dict(block=b, kernel_ms=ms, wall_ms=wall * 1e3,
                        chain_steps_per_s=nch * ns / (ms * 1e-3),
                        accept=float(o['accept_rate'].mean()),
                        e_per_n=float(o['sum_energy'][:, 0].mean() / ns / nop)))
    eng.close()
    return dict(config='C2 VMC N=50 1e5 chains M=50', blocks=out)


def dmc(nop, nw, nts, v0, modes=0, bins=0, label='', nblocks=3, dt=1e-3,
        profile=True):
    spec = model.Spec(v0, 1, 2, nop, nop, 0.25 * nop)
    eng = engine.Engine(spec)
    cap = int(nw * 1.25)
    dp = eng.dmc_params(dt, cap, nw, 0.5, 7, 0.0, float(nop),
                        ssf=(modes, True, nts) if modes else None,
                        density=(bins, True, nts) if bins else None)
    eng.dmc_init(dp, lattice_ini(nw, nop, 1))
    eng.dmc_run_block(nts)
    eng.set_profiling(profile)     # per-launch events rule out the block graph
    out = []
    den = np.zeros((nts, bins)) if bins else None
    ssf = np.zeros((nts, modes, 3)) if modes else None
    for b in range(nblocks):
        o = eng.dmc_run_block(nts, eval_estimators=bool(modes or bins), density=den, ssf=ssf)
        st = eng.last_block_stats()
        ws = float(o['num_walkers'].sum())
        out.append(dict(block=b, total_ms=st['total_ms'], step_kernel_ms=st['step_kernel_ms'],
                        walker_steps_per_s=ws / (st['total_ms'] * 1e-3),
                        step_kernel_share=st['step_kernel_ms'] / st['total_ms'],
                        e_per_n=float(o['energy'][-1] / o['weight'][-1] / nop)))
    eng.close()
    return dict(config=label, blocks=out)


if __name__ == '__main__':
    which = sys.argv[1:] or ['c2', 'c3', 'c5', 'c5e']
    if 'c2' in which:
        print(json.dumps(vmc_c2()), flush=True)
    if 'c2s' in which:      # short blocks: what an ncu replay can afford
        print(json.dumps(vmc_c2(ns=16, nblocks=3)), flush=True)
    if 'c3' in which:
        print(json.dumps(dmc(50, 10000, 512, 5 * PI ** 2, label='C3 DMC N=50 1e4 walkers')), flush=True)
    if 'c3g' in which:
        print(json.dumps(dmc(50, 10000, 512, 5 * PI ** 2, profile=False, label='C3 DMC N=50 1e4 walkers, block graph')), flush=True)
    if 'c3big' in which:
        print(json.dumps(dmc(50, 250000, 128, 5 * PI ** 2, label='DMC N=50 2.5e5 walkers (population scaling of C3)')), flush=True)
    if 'c5' in which:
        print(json.dumps(dmc(200, 31250, 64, 20 * PI ** 2, label='C5 DMC N=200 3.1e4 walkers/GPU, no estimators', dt=5e-4)), flush=True)
    if 'c5e' in which:
        print(json.dumps(dmc(200, 31250, 64, 20 * PI ** 2, modes=400, bins=6400,
                             label='C5 DMC N=200 3.1e4 walkers/GPU, S(k) M=400 + density B=6400 pure', dt=5e-4)), flush=True)
