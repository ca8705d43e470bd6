"""BASELINE configs[4] end to end through the sampler mirror: mrbp_qmc DMC,
N=200 deep lattice (V0 = 20 pi^2), 2.5e5 target walkers sharded over the ranks
of a torch.distributed job, pure S(k) (M=400) and pure density (B=6400)
estimators every step, a sweep over the time step.  One JSON line per time
step on rank 0.

    torchrun --nproc-per-node 8 scripts/run_config5.py [--target 250000]
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--target', type=int, default=250000)
    ap.add_argument('--nts', type=int, default=64)
    ap.add_argument('--burn', type=int, default=2)
    ap.add_argument('--blocks', type=int, default=4)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from phd_qmclib_b200 import dmc, model
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    torch.cuda.set_device(local)
    if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':
        del os.environ['NCCL_DEBUG']
    d = None
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        d = dist
    nop = 200
    spec = model.Spec(20 * math.pi ** 2, 1, 2, nop, nop, 0.25 * nop)
    rng = np.random.default_rng(1)
    # one particle per well, jittered (every rank builds the same global set)
    ini = np.zeros((args.target, 2, nop))
    ini[:, 0] = np.arange(nop)[None, :] + 0.25 \
        + 0.15 * (rng.random((args.target, nop)) - 0.5)
    for dt in (2e-3, 1e-3, 5e-4, 2.5e-4):
        smp = dmc.Sampling(spec, dt, int(1.25 * args.target), args.target,
                           num_walkers_control_factor=0.5, rng_seed=7,
                           dist=d, device=local,
                           ssf_est_spec=dmc.SSFEstSpec(400, True, args.nts),
                           density_est_spec=dmc.DensityEstSpec(6400, True,
                                                               args.nts))
        it = smp.blocks(smp.build_state(ini), args.nts, args.burn)
        e_sum = w_sum = ws = 0.0
        sk = np.zeros(400)
        den = np.zeros(6400)
        t0 = None
        for b in range(args.burn + args.blocks):
            if b == args.burn:
                torch.cuda.synchronize()
                if d is not None:
                    d.barrier()
                t0 = time.perf_counter()
            blk = next(it)
            if b < args.burn:
                continue
            ip = blk.iter_props
            e_sum += float(ip.energy.sum())
            w_sum += float(ip.weight.sum())
            ws += float(ip.num_walkers.sum())
            # pure estimators: the last step of the forward-walking window
            sk += np.asarray(blk.iter_ssf)[-1, :, 0] / float(ip.weight[-1])
            den += np.asarray(blk.iter_density)[-1, :, 0] \
                / float(ip.weight[-1])
        torch.cuda.synchronize()
        if d is not None:
            d.barrier()
        wall = time.perf_counter() - t0
        if rank == 0:
            sk /= args.blocks
            den /= args.blocks
            print(json.dumps({
                'config': 'BASELINE configs[4]: DMC N=200 V0=20pi^2, S(k) '
                          'M=400 pure + density B=6400 pure',
                'n_gpus': world, 'time_step': dt,
                'target_walkers': args.target, 'blocks': args.blocks,
                'steps_per_block': args.nts,
                'energy_per_particle': e_sum / w_sum / nop,
                'walker_steps_per_s_wall': ws / wall,
                'S_k_over_N_first_modes': (sk[1:6] / nop).tolist(),
                'density_sum_over_bins': float(den.sum()),
                'density_peak_to_mean': float(den.max() / den.mean()),
            }), flush=True)
        smp.engine.close()
    if d is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
