"""Step-kernel throughput at several (N, walkers), table path on/off
(development aid)."""
import math, os, sys
import numpy as np
sys.path.insert(0, '.')
from phd_qmclib_b200 import engine, model
PI = math.pi
for nop, nw in [tuple(map(int, a.split(":"))) for a in (sys.argv[1:] or ["50:250000", "100:125000", "200:31250", "20:200000"])]:
    spec = model.Spec(5 * PI ** 2, 1, 2, nop, nop, 0.25 * nop)
    eng = engine.Engine(spec)
    rng = np.random.default_rng(0)
    ini = np.zeros((nw, 2, nop))
    ini[:, 0] = np.arange(nop)[None, :] + 0.25 + 0.15 * (rng.random((nw, nop)) - 0.5)
    cap = int(nw * 1.25)
    dp = eng.dmc_params(6.25e-4, cap, nw, 0.5, 7, 0.0, float(nop))
    eng.dmc_init(dp, ini)
    eng.dmc_run_block(16)
    eng.set_profiling(True)
    out = eng.dmc_run_block(32)
    st = eng.last_block_stats()
    ws = float(out['num_walkers'].sum())
    print(f'tables={"off" if os.environ.get("QMCB_NO_TRIG_TABLES") else "on"} N={nop} W={nw}: '
          f'step-kernel {st["step_kernel_ms"] / 32:.4f} ms  {ws / (st["step_kernel_ms"] * 1e-3):.3e} ws/s (kernel only)', flush=True)
    eng.close()
