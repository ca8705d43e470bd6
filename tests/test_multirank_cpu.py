"""CPU, world_size 2 and 3 over gloo: the host-side logic of the sharded DMC
path -- the order-preserving rebalance plan (qmcb_rebalance_plan) executed
with torch.distributed send/recv on numpy walkers, and the global population
control recurrence fed by an all-reduce, against a single-rank run."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _plan(counts, rank):
    from phd_qmclib_b200 import _lib
    L = _lib.load()
    world = len(counts)
    cnt = np.asarray(counts, dtype=np.int64)
    send = np.zeros((world, 2), dtype=np.int64)
    recv = np.zeros((world, 2), dtype=np.int64)
    new = C.c_int64()
    rc = L.qmcb_rebalance_plan(_lib.ptr(cnt), world, rank, _lib.ptr(send),
                               _lib.ptr(recv), C.byref(new))
    assert rc == 0
    return send, recv, new.value


@pytest.mark.parametrize('counts', [[10, 2], [0, 9], [5, 5], [7, 0, 13],
                                    [1, 1, 1, 30], [100, 3, 50, 2, 77, 0, 9, 41]])
def test_plan_is_consistent(counts):
    """Every send has the matching receive, slabs stay contiguous, the new
    populations differ by at most one walker and the order is preserved."""
    world = len(counts)
    plans = [_plan(counts, r) for r in range(world)]
    total = sum(counts)
    news = [p[2] for p in plans]
    assert sum(news) == total and max(news) - min(news) <= 1
    # emulate the exchange on global walker ids
    off = np.concatenate([[0], np.cumsum(counts)])
    slabs = [np.arange(off[r], off[r + 1]) for r in range(world)]
    out = [np.full(news[r], -1, dtype=np.int64) for r in range(world)]
    for r in range(world):
        send, recv, _ = plans[r]
        for p in range(world):
            so, sn = send[p]
            ro, rn = plans[p][1][r]
            assert sn == rn
            out[p][ro:ro + rn] = slabs[r][so:so + sn]
    assert np.array_equal(np.concatenate(out), np.arange(total))


def _worker(rank, world, port, counts, nop, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(100 + rank)
        mine = counts[rank]
        confs = rng.random((mine, 2, nop))
        # tag each walker with its global id to check the order afterwards
        off = int(np.sum(counts[:rank]))
        confs[:, 0, 0] = off + np.arange(mine)
        # all-gather of the counts (qmcb_dmc_rebalance does it over NCCL)
        t = torch.zeros(world, dtype=torch.int64)
        t[rank] = mine
        dist.all_reduce(t)
        send, recv, new_n = _plan(t.tolist(), rank)
        new = np.zeros((new_n, 2, nop))
        reqs = []
        for p in range(world):
            so, sn = send[p]
            if sn and p != rank:
                reqs.append(dist.isend(
                    torch.from_numpy(confs[so:so + sn].copy()), p))
        for p in range(world):
            ro, rn = recv[p]
            if not rn:
                continue
            if p == rank:
                so = send[p][0]
                new[ro:ro + rn] = confs[so:so + rn]
            else:
                buf = torch.zeros((rn, 2, nop), dtype=torch.float64)
                dist.recv(buf, p)
                new[ro:ro + rn] = buf.numpy()
        for r in reqs:
            r.wait()
        # population control from the all-reduced sums: every rank derives
        # the same reference energy (qmc_base/dmc.py:758-771)
        s = torch.tensor([float(confs[:, 1].sum()), float(mine)],
                         dtype=torch.float64)
        dist.all_reduce(s)
        eref = s[0].item() / s[1].item() - 0.5 * np.log(s[1].item() / 64) / 1e-3
        q.put((rank, new[:, 0, 0].tolist(), float(new[:, 1].sum()), eref))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('counts', [[40, 6], [3, 31, 12]])
def test_rebalance_over_gloo(counts):
    world, nop = len(counts), 5
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker,
                         args=(r, world, port, counts, nop, q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids = np.concatenate([r[1] for r in res])
    assert np.array_equal(ids, np.arange(sum(counts)))        # order kept
    sizes = [len(r[1]) for r in res]
    assert max(sizes) - min(sizes) <= 1                        # balanced
    assert len({round(r[3], 9) for r in res}) == 1             # same E_ref


def _sampling_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from phd_qmclib_b200 import dmc, model
        spec = model.Spec(0, 1, 4, 16, 16, 4)
        smp = dmc.Sampling(spec, 1e-3, 101, 64, rng_seed=5, dist=dist,
                           ssf_est_spec=dmc.SSFEstSpec(4))
        assert (smp.world_size, smp.rank) == (world, rank)
        cap = smp.local_capacity
        share = -(-101 // world)
        assert cap == share + max(32, share // 4)
        assert smp.state_confs_shape == (cap, 2, 16)
        assert smp._engine_params().local_capacity == cap
        lo, hi = dmc.slab_bounds(64, world, rank)
        # estimator tables: every rank contributes its partial sums
        den = np.full((3, 5, 1), float(rank + 1))
        ssf = np.arange(3 * 4 * 3, dtype=np.float64).reshape(3, 4, 3) * (rank + 1)
        sums = np.array([10.0 * (rank + 1), float(hi - lo)])
        dmc.allreduce_sum(dist, [den, ssf])
        dmc.allreduce_sum(dist, [sums])
        q.put((rank, lo, hi, den.copy(), ssf.copy(), sums.copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_sampling_shards_and_reduces(world):
    """Host logic of the sharded dmc.Sampling: contiguous slabs that tile the
    ensemble in order, local capacity, and the per-block all-reduce of the
    estimator tables (gloo stands in for NCCL)."""
    port = _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_sampling_worker, args=(r, world, port, q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)],
                 key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tri = world * (world + 1) / 2
    edges = [0]
    for rank, lo, hi, den, ssf, sums in res:
        assert lo == edges[-1] and hi - lo in (64 // world, 64 // world + 1)
        edges.append(hi)
        assert np.all(den == tri)
        assert np.array_equal(
            ssf, np.arange(36, dtype=np.float64).reshape(3, 4, 3) * tri)
        assert sums[0] == 10.0 * tri and sums[1] == 64
    assert edges[-1] == 64


def test_sampling_sharded_needs_a_seed():
    from phd_qmclib_b200 import dmc, model

    class _Dist:
        @staticmethod
        def get_world_size():
            return 4

        @staticmethod
        def get_rank():
            return 1

    spec = model.Spec(0, 1, 4, 16, 16, 4)
    with pytest.raises(ValueError):
        dmc.Sampling(spec, 1e-3, 100, 64, dist=_Dist)
    smp = dmc.Sampling(spec, 1e-3, 100, 64, rng_seed=1, dist=_Dist)
    assert smp.local_capacity == 25 + 32
    assert smp.state_props_shape == (57,)
    assert [dmc.slab_bounds(10, 4, r) for r in range(4)] == [
        (0, 3), (3, 6), (6, 8), (8, 10)]


def _vmc_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from phd_qmclib_b200 import model, vmc
        spec = model.Spec(0, 1, 4, 16, 16, 4)
        smp = vmc.Sampling(spec, 0.25, rng_seed=5, dist=dist, chain_offset=7)
        q.put((rank, smp._chain_offset(10 + rank)))
    finally:
        dist.destroy_process_group()


def test_vmc_chains_are_numbered_across_ranks():
    """Rank r's chains follow those of the lower ranks in the RNG keying
    (10, 11 and 12 chains on ranks 0, 1, 2; base offset 7)."""
    world, port = 3, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_vmc_worker, args=(r, world, port, q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == {0: 7, 1: 17, 2: 28}
