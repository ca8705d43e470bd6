"""GPU, two ranks over NCCL (skipped on a single-GPU box): the sharded
dmc.Sampling mirror -- one process per GPU, global series identical on every
rank, estimator tables all-reduced per block, order-preserving rebalance
between blocks -- against a single-GPU run of the same ensemble."""
import os
import socket
import sys

import numpy as np
import pytest

from _blocking import ratio_mean_error
from specs import SPECS

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _ini(nop, size, n, seed):
    """One particle per cell, jittered inside the well: walkers of similar
    energy, so that the first branching steps do not collapse the ensemble
    onto a handful of ancestors (S(k) would then not self-average)."""
    rng = np.random.default_rng(seed)
    c = np.zeros((n, 2, nop))
    c[:, 0] = np.arange(nop)[None, :] + 0.25 + 0.2 * (rng.random((n, nop)) - 0.5)
    return c


NTS, NBLK, BURN, TARGET, WMAX, MODES, BINS = 64, 16, 8, 768, 1024, 6, 32


def _run(smp, ini):
    it = smp.blocks(smp.build_state(ini), NTS, BURN)
    out = dict(e=[], w=[], nw=[], ssf=[], den=[], local=[])
    for b in range(BURN + NBLK):
        blk = next(it)
        st = blk.last_state
        if b < BURN:
            continue
        out['e'].append(blk.iter_props.energy.copy())
        out['w'].append(blk.iter_props.weight.copy())
        out['nw'].append(blk.iter_props.num_walkers.copy())
        out['ssf'].append(np.asarray(blk.iter_ssf).copy())
        out['den'].append(np.asarray(blk.iter_density).copy())
        out['local'].append((int(st.num_walkers), float(st.weight),
                             int((~st.props.mask).sum())))
    return {k: np.array(v) for k, v in out.items()}


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    try:
        from phd_qmclib_b200 import dmc, model
        spec = model.Spec(**SPECS['lat_n50'])
        smp = dmc.Sampling(spec, 1e-3, WMAX, TARGET, rng_seed=9, dist=dist,
                           device=rank,
                           ssf_est_spec=dmc.SSFEstSpec(MODES, False, NTS),
                           density_est_spec=dmc.DensityEstSpec(BINS, False,
                                                               NTS))
        q.put((rank, _run(smp, _ini(50, 50.0, TARGET, 4))))
        smp.engine.close()
    finally:
        dist.destroy_process_group()


def _exact_worker(rank, world, port, q, energy_mode, every):
    """Raw engine, sharded: global ensemble of `n` walkers cut into ordered
    slabs; returns the (global) series of every block and the evolved slab."""
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    os.environ['QMCB_REBALANCE_EVERY'] = str(every)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    try:
        from conftest import golden
        from phd_qmclib_b200 import dmc, engine
        p = golden('model_ll_n16.npz')['params']
        ini = _exact_ini()
        lo, hi = dmc.slab_bounds(len(ini), world, rank)
        eng = engine.Engine((p[:12], p[12:19], p[19:]), device=rank)
        eng.comm_init_torch(dist, rank, world)
        cap = EX_WMAX // world + 150
        dp = eng.dmc_params(EX_DT, EX_WMAX, len(ini), 0.25, 11, 0.0, 16.0,
                            energy_mode=energy_mode, local_capacity=cap)
        eng.dmc_init(dp, ini[lo:hi])
        blocks, moved = [], 0
        for b in range(EX_BLOCKS):
            blocks.append(eng.dmc_run_block(EX_NTS))
            if b == 1:
                moved += eng.dmc_rebalance()
        nx = eng.dmc_get_next()
        hits = int(eng.dmc_scalars().capacity_hits)
        q.put((rank, blocks, nx['confs'], nx['weight'], nx['energy'], moved,
               hits))
        eng.close()
    finally:
        dist.destroy_process_group()


EX_WMAX, EX_DT, EX_NTS, EX_BLOCKS = 1200, 4e-3, 24, 3


def _exact_ini():
    rng = np.random.default_rng(77)
    ini = np.zeros((900, 2, 16))
    ini[:, 0] = rng.random((900, 16)) * 16
    return ini


@pytest.mark.parametrize('energy_mode,every', [(0, 0), (0, 8), (1, 8)])
def test_two_ranks_reproduce_the_single_rank_oracle(oracle, energy_mode,
                                                    every):
    """The sharded engine against the SERIAL oracle on the whole ensemble:
    the RNG is keyed by global positions (parents) and clone indices, the
    stale-slot energies of quirk Q1 live in one global per-position array
    replicated on the ranks, and the rebalance (between blocks and, with
    `every` > 0, inside them) preserves the order -- so two ranks must walk
    the very same ensemble as one: identical walker counts at every step,
    series to rounding, and the concatenated slabs equal to the oracle's
    population."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    import torch.multiprocessing as mp
    from conftest import golden
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_exact_worker,
                         args=(r, world, port, q, energy_mode, every))
             for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)],
                 key=lambda x: x[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    p = golden('model_ll_n16.npz')['params']
    ini = _exact_ini()
    st = oracle.DMCState(p, ini, EX_WMAX)
    for b in range(EX_BLOCKS):
        a = st.run_block(11, EX_DT, len(ini), 0.25, EX_NTS, 0.0, 16.0,
                         energy_mode=energy_mode)
        for r in range(world):
            got = res[r][1][b]
            assert np.array_equal(a['num_walkers'], got['num_walkers']), b
            for k in ('energy', 'weight', 'ref_energy', 'accum_energy'):
                assert np.max(np.abs(got[k] - a[k]) / np.abs(a[k])) < 1e-10, k
    assert all(r[6] == 0 for r in res)              # no slab hit its capacity
    nw = st.num_walkers
    confs = np.concatenate([r[2] for r in res])
    weight = np.concatenate([r[3] for r in res])
    energy = np.concatenate([r[4] for r in res])
    assert confs.shape[0] == nw
    assert np.allclose(confs[:, 0], st.prev['confs'][:nw, 0], rtol=0,
                       atol=1e-9)
    assert np.allclose(weight, st.prev['weight'][:nw], rtol=1e-9)
    assert np.allclose(energy, st.prev['energy'][:nw], rtol=1e-9, atol=1e-9)
    # the populations did move between the ranks
    assert sum(r[5] for r in res) > 0
    assert a['num_walkers'].min() != a['num_walkers'].max()


def test_sharded_sampling_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    import torch.multiprocessing as mp
    from phd_qmclib_b200 import dmc, model
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    a, b = res[0], res[1]
    # the per-step series and the all-reduced estimator tables are global
    for k in ('e', 'w', 'nw', 'ssf', 'den'):
        assert np.array_equal(a[k], b[k]), k
    # local slabs add up to the global population of the last step
    nw_last = a['nw'][:, -1].astype(np.int64)
    assert np.array_equal(a['local'][:, 0] + b['local'][:, 0], nw_last)
    assert np.array_equal(a['local'][:, 2], a['local'][:, 0])
    assert np.allclose(a['local'][:, 1], a['w'][:, -1])
    # the rebalance keeps the slabs within one walker of each other at the
    # start of a block; after NTS steps they have drifted only a little
    assert np.all(np.abs(a['local'][:, 0] - b['local'][:, 0]) < 0.2 * TARGET)
    # mixed S(k), k = 0: N^2 per live walker; density: N per live walker
    assert np.allclose(a['ssf'][:, :, 0, 0], 50.0 ** 2 * a['w'])
    # density: N counts per live walker in the first two steps of a block
    # (later steps accumulate across parity buffers, the reference's quirk
    # Q3, pinned step by step in test_sampling_gpu.py)
    dsum = a['den'].sum(axis=2)[..., 0]
    assert np.allclose(dsum[:, :2], 50.0 * a['w'][:, :2])
    # same physics as one GPU holding the whole ensemble
    spec = model.Spec(**SPECS['lat_n50'])
    one = dmc.Sampling(spec, 1e-3, WMAX, TARGET, rng_seed=9,
                       ssf_est_spec=dmc.SSFEstSpec(MODES, False, NTS),
                       density_est_spec=dmc.DensityEstSpec(BINS, False, NTS))
    c = _run(one, _ini(50, 50.0, TARGET, 4))
    one.engine.close()
    e2, s2 = ratio_mean_error(a['e'].sum(axis=1), a['w'].sum(axis=1))
    e1, s1 = ratio_mean_error(c['e'].sum(axis=1), c['w'].sum(axis=1))
    assert abs(e2 - e1) < 5 * np.hypot(s1, s2), (e2 / 50, e1 / 50)
    for m in range(1, MODES):
        k2, d2 = ratio_mean_error(a['ssf'][:, :, m, 0].sum(axis=1),
                                  a['w'].sum(axis=1))
        k1, d1 = ratio_mean_error(c['ssf'][:, :, m, 0].sum(axis=1),
                                  c['w'].sum(axis=1))
        print(f'S(k) mode {m}: two GPUs {k2:.4f}({d2:.4f}) one {k1:.4f}({d1:.4f})')
        assert abs(k2 - k1) < 6 * np.hypot(d1, d2), (m, k2, k1, d1, d2)
