"""Development aid: z-scores of the pure density bins, engine run vs engine
run and engine vs the frozen reference run (which of the two carries an rms
above 1 tells a bias from under-estimated errors)."""
import sys
from itertools import islice

import numpy as np

sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
from _blocking import ratio_mean_error  # noqa: E402
from phd_qmclib_b200 import dmc  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'lat_n100'
g = np.load(f'tests/golden/dmc_stat_pure_{name}.npz')
p = g['params']
nop = int(p[3])
nts, nblocks, burn = int(g['nts']), int(g['nblocks']), int(g['burn'])
M, B = int(g['num_modes']), int(g['num_bins'])


class Spec:
    params, obf_params, tbf_params = p[:12], p[12:19], p[19:]
    boson_number, supercell_size = nop, float(p[4])
    boundaries = (0.0, float(p[4]))
    sys_conf_shape = (2, nop)


def run(seed):
    smp = dmc.Sampling(Spec, float(g['time_step']), int(g['max_num_walkers']),
                       int(g['n_target']),
                       num_walkers_control_factor=float(g['nwc_factor']),
                       rng_seed=seed,
                       ssf_est_spec=dmc.SSFEstSpec(M, True, nts),
                       density_est_spec=dmc.DensityEstSpec(B, True, nts))
    it = smp.blocks(smp.build_state(g['ini_confs']), nts, burn)
    for _ in islice(it, burn):
        pass
    d, n = [], []
    for _, blk in zip(range(nblocks), it):
        d.append(np.asarray(blk.iter_density)[nts - 1, :, 0].copy())
        n.append(float(blk.iter_props.num_walkers[nts - 1]))
    smp.engine.close()
    return np.array(d), np.array(n)


def zs(a, an, b, bn):
    out = []
    for k in range(B):
        x, dx = ratio_mean_error(a[:, k], an)
        y, dy = ratio_mean_error(b[:, k], bn)
        out.append((x - y) / np.hypot(dx, dy))
    out = np.array(out)
    return f'max {np.abs(out).max():.2f} rms {np.sqrt(np.mean(out**2)):.2f}'


runs = [run(s) for s in (99, 100, 101, 102)]
ref = (g['block_density_last'], g['block_walkers_last'])
for i in range(4):
    print(f'engine {i} vs reference:', zs(*runs[i], *ref))
for i in range(4):
    for j in range(i + 1, 4):
        print(f'engine {i} vs engine {j}:', zs(*runs[i], *runs[j]))
