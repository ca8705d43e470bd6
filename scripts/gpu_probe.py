"""Quick on-GPU probe: fp64 peak, fast_rcp accuracy via model parity, DMC
throughput at a few sizes.  Development aid, not the bench."""
import ctypes as C
import math
import sys
import time

import numpy as np

sys.path.insert(0, '.')
from phd_qmclib_b200 import _lib, engine, model  # noqa: E402

L = _lib.load()
tf, ms = C.c_double(), C.c_double()
rc = L.qmcb_measure_fp64_peak(0, C.byref(tf), C.byref(ms))
print(f'fp64 DFMA peak: rc={rc} {tf.value:.2f} TFLOP/s ({ms.value:.3f} ms)')

PI = math.pi
for nop, nw in [(100, 20000), (100, 100000), (50, 10000), (50, 100000),
                (200, 20000), (20, 100000)]:
    spec = model.Spec(5 * PI ** 2, 1, 2, nop, nop, 0.25 * nop)
    eng = engine.Engine(spec)
    rng = np.random.default_rng(0)
    ini = np.zeros((nw, 2, nop))
    ini[:, 0] = rng.random((nw, nop)) * nop
    cap = int(nw * 1.25)
    dp = eng.dmc_params(6.25e-4, cap, nw, 0.5, 7, 0.0, float(nop))
    t0 = time.time()
    eng.dmc_init(dp, ini)
    t_init = time.time() - t0
    eng.dmc_run_block(8)          # warm-up / equilibrate a little
    eng.set_profiling(True)
    nts = 16
    out = eng.dmc_run_block(nts)
    st = eng.last_block_stats()
    ws = float(out['num_walkers'].sum())
    F = 58 * nop * (nop - 1) / 2 + 113 * nop + 35
    print(f'N={nop} W={nw}: init {t_init:.2f}s  block {st["total_ms"]:.2f} ms'
          f' step-kernel {st["step_kernel_ms"]:.2f} ms  '
          f'{ws / (st["total_ms"] * 1e-3):.3e} ws/s  '
          f'alg {ws * F / (st["step_kernel_ms"] * 1e-3) / 1e12:.2f} TF  '
          f'nw[-1]={int(out["num_walkers"][-1])} E/N='
          f'{out["energy"][-1] / out["weight"][-1] / nop:.4f}')
    eng.close()
