"""Statistical parity of the SHARDED engine against the frozen runs of the
live reference (tests/golden/dmc_stat_*.npz), at any world size:

    python scripts/stat_parity_multirank.py                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
        --master-addr 127.0.0.1 scripts/stat_parity_multirank.py

Same model, time step, global population (512 target / 640 slots, i.e. 64
walkers per rank on 8 GPUs -- the worst case for anything that depends on the
slab boundaries) and block structure as the reference run; prints the z-score
of the DMC energy and of the mixed S(k) estimator against the reference with
combined blocked errors (the criterion of tests/test_sampling_gpu.py).
"""
import json
import os
import sys
from itertools import islice

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from _blocking import ratio_mean_error  # noqa: E402


def main():
    import torch
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl',
                                device_id=torch.device('cuda', local))
    from phd_qmclib_b200 import dmc
    mult = int(os.environ.get('QMCB_STAT_MULT', '4'))
    energy_mode = int(os.environ.get('QMCB_STAT_ENERGY_MODE', '0'))
    out = []
    for name in ('ll_n16', 'lat_n16'):
        g = np.load(os.path.join(ROOT, 'tests', 'golden',
                                 f'dmc_stat_{name}.npz'))
        p = g['params']
        nop = int(p[3])
        nts, nblocks, burn = int(g['nts']), int(g['nblocks']), int(g['burn'])

        class _Spec:
            params, obf_params, tbf_params = p[:12], p[12:19], p[19:]
            boson_number, supercell_size = nop, float(p[4])
            boundaries = (0.0, float(p[4]))
            sys_conf_shape = (2, nop)

        num_modes = int(g['num_modes'])
        smp = dmc.Sampling(_Spec, float(g['time_step']),
                           int(g['max_num_walkers']), int(g['n_target']),
                           num_walkers_control_factor=float(g['nwc_factor']),
                           rng_seed=2024, dist=dist, device=local,
                           energy_mode=energy_mode,
                           ssf_est_spec=dmc.SSFEstSpec(num_modes, False, nts))
        it = smp.blocks(smp.build_state(g['ini_confs']), nts, burn)
        for _ in islice(it, burn):
            pass
        e_sum, w_sum, s_sum = [], [], []
        for _, blk in zip(range(mult * nblocks), it):
            e_sum.append(blk.iter_props.energy.sum())
            w_sum.append(blk.iter_props.weight.sum())
            s_sum.append(np.asarray(blk.iter_ssf)[:, :, 0].sum(axis=0))
        hits = int(smp.engine.dmc_scalars().capacity_hits)
        smp.engine.close()
        e_ref, err_ref = ratio_mean_error(g['block_energy'],
                                          g['block_weight'])
        err_ref = max(err_ref, float(g['ref_energy_err']))
        e_eng, err_eng = ratio_mean_error(e_sum, w_sum)
        sig = float(np.hypot(err_ref, err_eng))
        rec = dict(name=name, world=world, energy_mode=energy_mode,
                   e_per_n_engine=e_eng / nop, e_per_n_reference=e_ref / nop,
                   err_engine=err_eng / nop, err_reference=err_ref / nop,
                   z_energy=(e_eng - e_ref) / sig, capacity_hits_rank0=hits,
                   blocks=mult * nblocks, z_ssf=[])
        s_sum = np.array(s_sum)
        for m in range(1, num_modes):
            sk_ref, sk_err_ref = ratio_mean_error(g['block_ssf'][:, m, 0],
                                                  g['block_weight'])
            sk_err_ref = max(sk_err_ref, float(g['ref_sk_err'][m]))
            sk_eng, sk_err_eng = ratio_mean_error(s_sum[:, m], w_sum)
            rec['z_ssf'].append(float((sk_eng - sk_ref)
                                      / np.hypot(sk_err_ref, sk_err_eng)))
        out.append(rec)
        if rank == 0:
            print(json.dumps(rec), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
