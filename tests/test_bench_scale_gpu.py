"""GPU parity at the populations the bench runs: capacities far beyond one
branching tile (BR_TILE = 1024 slots), so that the multi-CTA form of
`sync_branching_spec` (qmc_base/dmc.py:614-655) -- per-CTA counts, the offset
of a CTA from the totals of the CTAs before it, the truncation
at capacity across CTA boundaries and the fixed-order sum of the cloned
parents' energies (state_energy, qmc_base/dmc.py:759-760) -- is compared with
the serial oracle on the same Philox streams.  Bit-equal walker counts and
cloning tables; per-step series to 1e-10."""
import numpy as np
import pytest

from conftest import golden, maxnorm_err, rel_err

pytestmark = pytest.mark.gpu

BR_TILE = 1024      # phd_qmclib_b200/csrc/qmcb_kernels.cuh


def _spec(p):
    return (p[:12], p[12:19], p[19:])


def _ini(rng, n, nop, size):
    ini = np.zeros((n, 2, nop))
    ini[:, 0] = rng.random((n, nop)) * size
    return ini


def _check_state(eng, st):
    """Cloning table, mask, positions and energies of the yielded state, then
    the evolved population and the stale-slot energies (quirk Q1)."""
    s = eng.dmc_get_state()
    nw = st.num_walkers
    assert int(s['scalars'].num_walkers) == nw
    assert np.array_equal(s['cloning_ref'][:nw], st.ref[:nw])
    assert np.array_equal(s['mask'], st.act['mask'])
    assert np.allclose(s['confs'][:nw, 0], st.act['confs'][:nw, 0],
                       rtol=0, atol=1e-9)
    assert rel_err(s['energy'][:nw], st.act['energy'][:nw]) < 1e-9
    nx = eng.dmc_get_next()
    assert np.allclose(nx['confs'][:, 0], st.prev['confs'][:nw, 0], rtol=0,
                       atol=1e-9)
    assert maxnorm_err(nx['confs'][:, 1], st.prev['confs'][:nw, 1]) < 1e-8
    assert rel_err(nx['weight'], st.prev['weight'][:nw]) < 1e-9
    assert np.allclose(nx['slot_energy'], st.act['energy'], rtol=1e-9,
                       atol=1e-9)


@pytest.mark.parametrize('energy_mode', [0, 1])
def test_dmc_n100_8192_walkers_vs_oracle(oracle, energy_mode):
    """BASELINE configs[3] model (N=100) at 8192 target / 10240 slots: ten
    branching CTAs.  Time step large enough that every step has births and
    deaths in every CTA."""
    from phd_qmclib_b200 import engine
    p = golden('model_lat_n100.npz')['params']
    nop, size = int(p[3]), float(p[4])
    target, wmax, nts, dt, seed = 8192, 10240, 6, 2e-3, 1234
    ini = _ini(np.random.default_rng(100), target, nop, size)
    st = oracle.DMCState(p, ini, wmax)
    eng = engine.Engine(_spec(p))
    dp = eng.dmc_params(dt, wmax, target, 0.5, seed, 0.0, size,
                        energy_mode=energy_mode)
    eng.dmc_init(dp, ini)
    births = deaths = 0
    for _ in range(2):
        a = st.run_block(seed, dt, target, 0.5, nts, 0.0, size,
                         energy_mode=energy_mode)
        b = eng.dmc_run_block(nts)
        assert np.array_equal(a['num_walkers'], b['num_walkers'])
        for k in ('energy', 'weight', 'ref_energy', 'accum_energy'):
            assert rel_err(b[k], a[k]) < 1e-10, k
        nw = st.num_walkers
        cnt = np.bincount(st.ref[:nw], minlength=wmax)
        births += int((cnt > 1).sum())
        deaths += int((cnt[:int(st.ref[nw - 1]) + 1] == 0).sum())
    assert births > 10 and deaths > 10      # the test did branch
    assert st.num_walkers > 5 * BR_TILE
    _check_state(eng, st)
    eng.close()


@pytest.mark.parametrize('n_ini,wmax', [(3000, 3500), (2040, 2 * BR_TILE + 1),
                                        (5000, 5 * BR_TILE)])
def test_capacity_hit_inside_a_later_cta(oracle, n_ini, wmax):
    """Quirk Q8 (qmc_base/dmc.py:636-651): the reference fills slots in
    parent order and drops everything past the capacity.  A reference energy
    far too high for the first step makes the population overflow; the cut
    falls among the parents of a branching CTA with index >= 1 or 2 (for
    wmax = 2049 one slot into the third tile of children; for wmax = 5120
    exactly on a tile boundary).  Compared step by step."""
    from phd_qmclib_b200 import engine
    p = golden('model_ll_n16.npz')['params']
    nop, size = int(p[3]), float(p[4])
    ini = _ini(np.random.default_rng(n_ini), n_ini, nop, size)
    st = oracle.DMCState(p, ini, wmax, ref_energy=60.0)
    eng = engine.Engine(_spec(p))
    dp = eng.dmc_params(5e-3, wmax, n_ini, 0.0, 9, 0.0, size)
    eng.dmc_init(dp, ini, ref_energy=60.0)
    cut_parents = []
    for _ in range(4):
        a = st.run_block(9, 5e-3, n_ini, 0.0, 1, 0.0, size)
        b = eng.dmc_run_block(1)
        assert np.array_equal(a['num_walkers'], b['num_walkers'])
        for k in ('energy', 'weight', 'ref_energy', 'accum_energy'):
            assert rel_err(b[k], a[k]) < 1e-10, k
        nw = st.num_walkers
        s = eng.dmc_get_state()
        assert np.array_equal(s['cloning_ref'][:nw], st.ref[:nw])
        if nw == wmax:
            cut_parents.append(int(st.ref[nw - 1]))
    # the population did hit the capacity, and the last parent that got a
    # slot sat in a later branching CTA
    assert cut_parents and max(cut_parents) >= min(2 * BR_TILE, n_ini) - BR_TILE
    assert max(cut_parents) // BR_TILE >= (2 if n_ini > 2 * BR_TILE else 1)
    assert eng.dmc_scalars().capacity_hits >= 1
    _check_state(eng, st)
    eng.close()


def test_dmc_config3_n50_10000_walkers_vs_oracle(oracle):
    """BASELINE configs[2]: N=50, 1e4 target walkers, capacity 1.25e4, dt=1e-3,
    kappa=0.5 (Proc default) -- thirteen branching CTAs, the CUDA-graph block
    path (no estimators, one rank)."""
    from phd_qmclib_b200 import engine
    p = golden('model_lat_n50.npz')['params']
    nop, size = int(p[3]), float(p[4])
    target, wmax, nts, dt, seed = 10000, 12500, 8, 1e-3, 77
    ini = _ini(np.random.default_rng(50), target, nop, size)
    st = oracle.DMCState(p, ini, wmax)
    eng = engine.Engine(_spec(p))
    dp = eng.dmc_params(dt, wmax, target, 0.5, seed, 0.0, size)
    eng.dmc_init(dp, ini)
    for _ in range(3):
        a = st.run_block(seed, dt, target, 0.5, nts, 0.0, size)
        b = eng.dmc_run_block(nts)
        assert np.array_equal(a['num_walkers'], b['num_walkers'])
        for k in ('energy', 'weight', 'ref_energy', 'accum_energy'):
            assert rel_err(b[k], a[k]) < 1e-10, k
    _check_state(eng, st)
    eng.close()


def test_dmc_config5_n200_estimators_vs_oracle(oracle):
    """BASELINE configs[4] shape: N=200 deep lattice, 4096 target walkers,
    S(k) with M=400 modes and density with B=6400 bins, both pure (forward
    walking), estimator series of the second block against the oracle."""
    from phd_qmclib_b200 import engine
    p = golden('model_deep_n200.npz')['params']
    nop, size = int(p[3]), float(p[4])
    target, wmax, nts, dt, seed = 4096, 5120, 3, 1e-3, 5
    modes, bins = 400, 6400
    ini = _ini(np.random.default_rng(200), target, nop, size)
    st = oracle.DMCState(p, ini, wmax)
    ssf = dict(num=modes, pure=True, pfw=nts, iter=np.zeros((nts, modes, 3)),
               aux=np.zeros((2, wmax, modes, 3)))
    den = dict(num=bins, pure=True, pfw=nts, iter=np.zeros((nts, bins)),
               aux=np.zeros((2, wmax, bins)))
    eng = engine.Engine(_spec(p))
    dp = eng.dmc_params(dt, wmax, target, 0.5, seed, 0.0, size,
                        ssf=(modes, True, nts), density=(bins, True, nts))
    eng.dmc_init(dp, ini)
    for blk in range(2):
        est = blk > 0
        for d in (ssf, den):
            d['iter'][:] = 0
            d['aux'][:] = 0
        a = st.run_block(seed, dt, target, 0.5, nts, 0.0, size, eval_est=est,
                         ssf=ssf, density=den)
        e_den, e_ssf = np.zeros((nts, bins)), np.zeros((nts, modes, 3))
        b = eng.dmc_run_block(nts, eval_estimators=est, density=e_den,
                              ssf=e_ssf)
        assert np.array_equal(a['num_walkers'], b['num_walkers'])
        for k in ('energy', 'weight', 'ref_energy', 'accum_energy'):
            assert rel_err(b[k], a[k]) < 1e-10, k
        if est:
            assert np.allclose(e_den, den['iter'], rtol=1e-13, atol=1e-13)
            assert np.max(np.abs(e_ssf - ssf['iter'])) \
                < 1e-11 * np.max(np.abs(ssf['iter']))
    eng.close()


def test_branching_table_matches_serial_prefix_at_scale(oracle):
    """The cloning table alone, at 1.5e5 slots (147 branching CTAs, the bench
    shard): run one engine step from explicit weights and compare the whole
    child->parent table with the oracle's serial loop on the same uniforms."""
    from phd_qmclib_b200 import engine
    from phd_qmclib_b200._lib import StateScalars
    p = golden('model_ll_n16.npz')['params']
    nop, size = int(p[3]), float(p[4])
    n, wmax, seed = 125000, 150000, 4242
    rng = np.random.default_rng(9)
    confs = _ini(rng, n, nop, size)
    eng = engine.Engine(_spec(p))
    dp = eng.dmc_params(1e-3, wmax, n, 0.5, seed, 0.0, size)
    ev = eng.model_eval(confs, want=('energy', 'drift'))
    confs[:, 1] = ev['drift']
    weight = np.exp(rng.normal(0.0, 0.35, size=n))
    weight[::97] = 0.0                       # certain deaths
    weight[5::1013] = 7.3                    # many children of one parent
    sc = StateScalars()
    sc.num_walkers, sc.max_num_walkers = n, wmax
    sc.ref_energy = float(ev['energy'].mean())
    sc.weight, sc.energy = float(n), float(ev['energy'].sum())
    eng.dmc_set_state(dp, confs, ev['energy'], weight, sc)
    b = eng.dmc_run_block(1)
    s = eng.dmc_get_state()
    # the oracle's serial loop on the same per-slot Philox uniforms
    u = np.array([oracle.rng_uniform2(seed, i, 0, 0, 0)[0] for i in range(n)])
    nw, ref = oracle.branch(weight, n, wmax, u)
    assert int(b['num_walkers'][0]) == nw
    assert np.array_equal(s['cloning_ref'][:nw], ref[:nw])
    want_e = float(ev['energy'][ref[:nw]].sum())
    assert abs(b['energy'][0] - want_e) < 1e-11 * abs(want_e)
    eng.close()


def _vmc_vs_oracle(oracle, spec_kwargs, nch, ns, nblocks, modes, seed):
    from phd_qmclib_b200 import engine, model
    spec = model.Spec(**spec_kwargs)
    p = model.param_block(spec)
    nop, size = spec.boson_number, float(spec.supercell_size)
    spread = 0.25 * spec.well_width
    ini = _ini(np.random.default_rng(seed), nch, nop, size)
    cur = ini.copy()
    ln = oracle.model_eval(p, cur, want=('lnpsi',))['lnpsi']
    eprev, sprev = np.zeros(nch), np.zeros((nch, modes, 3))
    eng = engine.Engine(spec)
    eng.vmc_init(ini, spread, seed, 0.0, size, ssf_num_modes=modes)
    step0 = 0
    for b in range(nblocks):
        first = b == 0
        a = oracle.vmc_block(p, seed, spread, 0.0, size, cur, ln, eprev,
                             sprev, modes, ns, step0, first)
        step0 += ns - (1 if first else 0)
        o = eng.vmc_run_block(ns, series=True, sums=True)
        assert np.array_equal(o['move_stat'], a['stat'])
        assert np.max(np.abs(o['lnpsi'] - a['lnpsi'])
                      / np.maximum(np.abs(a['lnpsi']), 1.0)) < 1e-11
        assert np.max(np.abs(o['energy'] - a['energy'])
                      / np.maximum(np.abs(a['energy']), 1.0)) < 1e-11
        assert np.max(np.abs(o['ssf'] - a['ssf'])) < 1e-9
        assert np.allclose(o['sum_ssf'], a['ssf'].sum(axis=1), rtol=1e-9,
                           atol=1e-7)
    confs, lnpsi = eng.vmc_get_state()
    assert np.allclose(confs[:, 0], cur[:, 0], rtol=0, atol=1e-12)
    eng.close()
    return a


def test_vmc_config1_n20_single_chain_vs_oracle(oracle):
    """BASELINE configs[0] (the reference's own CPU case): N = 20 bosons,
    V0 = 5 E_R, g = 4, one Metropolis chain, move spread a quarter of the
    well, M = 20 modes (tests/mrbp_qmc/test_vmc_exec_proc.py:8-26): 3 blocks
    of 2048 steps, accept/reject sequence identical to the oracle's."""
    a = _vmc_vs_oracle(oracle, dict(
        lattice_depth=5 * np.pi ** 2, lattice_ratio=1, interaction_strength=4,
        boson_number=20, supercell_size=20, tbf_contact_cutoff=5), 1, 2048, 3,
        20, 1)
    assert 0.2 < a['accept_rate'][0] < 0.95


def test_vmc_config2_n50_chains_vs_oracle(oracle):
    """BASELINE configs[1] shape: N = 50 (a ragged last particle block), a
    batch of chains with energy and S(k) M = 50, table path of the block
    kernel."""
    _vmc_vs_oracle(oracle, dict(
        lattice_depth=5 * np.pi ** 2, lattice_ratio=1, interaction_strength=4,
        boson_number=50, supercell_size=50, tbf_contact_cutoff=12.5), 1500,
        24, 2, 50, 3)


def test_vmc_config2_full_size_sums_only(oracle):
    """BASELINE configs[1] at its full size (1e5 chains, N = 50, M = 50) on
    the sums-only path the bench times (no per-step series; S(k) block sums
    accumulated lazily on the device): size-independent identities for every
    chain, and a slice of chains from the middle of the batch against the
    oracle (same Philox keys: the chain index is part of the key)."""
    from phd_qmclib_b200 import engine, model
    spec = model.Spec(lattice_depth=5 * np.pi ** 2, lattice_ratio=1,
                      interaction_strength=4, boson_number=50,
                      supercell_size=50, tbf_contact_cutoff=12.5)
    p = model.param_block(spec)
    nop, size, M = 50, 50.0, 50
    nch, ns, seed = 100000, 48, 9
    spread = 0.25 * spec.well_width
    rng = np.random.default_rng(4)
    ini = np.zeros((nch, 2, nop))
    ini[:, 0] = (np.arange(nop)[None, :] + 0.25
                 + 0.15 * (rng.random((nch, nop)) - 0.5))
    k0, nk = 61234, 40                  # straddles CTA boundaries (G = 19)
    cur = ini[k0:k0 + nk].copy()
    ln = oracle.model_eval(p, cur, want=('lnpsi',))['lnpsi']
    eprev, sprev = np.zeros(nk), np.zeros((nk, M, 3))
    eng = engine.Engine(spec)
    eng.vmc_init(ini, spread, seed, 0.0, size, ssf_num_modes=M)
    step0 = 0
    for b in range(2):
        first = b == 0
        a = oracle.vmc_block(p, seed, spread, 0.0, size, cur, ln, eprev,
                             sprev, M, ns, step0, first, chain_offset=k0)
        step0 += ns - (1 if first else 0)
        o = eng.vmc_run_block(ns, series=False, sums=True)
        # every step adds |rho_0|^2 = N^2 and rho_0 = N, whatever the moves
        assert np.array_equal(o['sum_ssf'][:, 0, 0],
                              np.full(nch, float(ns * nop * nop)))
        assert np.array_equal(o['sum_ssf'][:, 0, 1],
                              np.full(nch, float(ns * nop)))
        assert np.all(o['sum_ssf'][:, 0, 2] == 0.0)
        assert np.all(o['sum_ssf'][:, :, 0] >= 0.0)
        assert np.all(np.isfinite(o['sum_energy']))
        # Cauchy-Schwarz of the two energy sums
        assert np.all(o['sum_energy'][:, 0] ** 2
                      <= ns * o['sum_energy'][:, 1] * (1 + 1e-12))
        assert 0.3 < o['accept_rate'].mean() < 0.7
        sl = slice(k0, k0 + nk)
        assert np.allclose(o['accept_rate'][sl], a['accept_rate'], rtol=0,
                           atol=1e-15)
        assert np.allclose(o['sum_energy'][sl, 0], a['energy'].sum(axis=1),
                           rtol=1e-11)
        assert np.allclose(o['sum_ssf'][sl], a['ssf'].sum(axis=1), rtol=1e-9,
                           atol=1e-7)
    confs, lnpsi = eng.vmc_get_state()
    assert np.allclose(confs[k0:k0 + nk, 0], cur[:, 0], rtol=0, atol=1e-12)
    assert rel_err(lnpsi[k0:k0 + nk], ln) < 1e-11
    eng.close()


@pytest.mark.parametrize('energy_mode', [0, 1])
def test_sharded_code_path_on_one_rank_vs_oracle(oracle, energy_mode):
    """The multi-rank machinery on a communicator of ONE rank (what a
    single-GPU box can run): per-step pack -> all-reduce -> all-gather ->
    population control on the second stream, the weights from the global
    per-position stale-energy array in their own kernel behind the step
    kernel, that array's update, its export/import across a restart and the
    rebalance entry points -- against the serial oracle, with a population
    spanning several branching CTAs.  (Two real ranks against the oracle:
    tests/test_multigpu_gpu.py.)"""
    import torch  # noqa: F401  (maps the NCCL the engine dlopens)
    from phd_qmclib_b200 import engine
    p = golden('model_ll_n16.npz')['params']
    nop, size = int(p[3]), float(p[4])
    n, wmax, nts, dt, seed = 3000, 4096, 10, 4e-3, 17
    ini = _ini(np.random.default_rng(23), n, nop, size)
    st = oracle.DMCState(p, ini, wmax)
    eng = engine.Engine(_spec(p))
    eng.comm_init(engine.comm_unique_id(), 1, 0)
    dp = eng.dmc_params(dt, wmax, n, 0.25, seed, 0.0, size,
                        energy_mode=energy_mode, local_capacity=wmax)
    eng.dmc_init(dp, ini)
    for blk in range(3):
        a = st.run_block(seed, dt, n, 0.25, nts, 0.0, size,
                         energy_mode=energy_mode)
        b = eng.dmc_run_block(nts)
        assert np.array_equal(a['num_walkers'], b['num_walkers'])
        for k in ('energy', 'weight', 'ref_energy', 'accum_energy'):
            assert rel_err(b[k], a[k]) < 1e-10, k
        assert eng.dmc_rebalance() == 0         # one rank: nothing to move
        if blk == 1:
            # restart on a fresh handle, through the exported view of the
            # global array
            nx = eng.dmc_get_next()
            eng.close()
            eng = engine.Engine(_spec(p))
            eng.comm_init(engine.comm_unique_id(), 1, 0)
            eng.dmc_set_state(dp, nx['confs'], nx['energy'], nx['weight'],
                              nx['scalars'], slot_energy=nx['slot_energy'])
    # per step: 3 + pack + finalize (+ the global array's update); per block:
    # begin (+ the weights of the last step and that step's update)
    assert eng.last_block_stats()['launches'] == 1 + nts * (
        6 if energy_mode == 0 else 5) + (1 if energy_mode == 0 else 0)
    _check_state(eng, st)
    eng.close()
