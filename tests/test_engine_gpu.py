"""GPU parity tests: the CUDA engine, through the C ABI, against (a) outputs
of the live reference frozen in tests/golden/ and (b) the C oracle on the same
seeded inputs.  Tolerance: 1e-12 relative (BASELINE.json north_star); drift in
max-norm."""
import numpy as np
import pytest

from conftest import (conditioned_rel_err, golden, golden_names, maxnorm_err,
                      rel_err, scaled_err)

pytestmark = pytest.mark.gpu

TOL = 1e-12


@pytest.fixture(scope='module')
def eng_mod():
    from phd_qmclib_b200 import engine
    return engine


@pytest.mark.parametrize('name', golden_names('model_'))
def test_model_eval_vs_reference(eng_mod, name):
    g = golden(name)
    with eng_mod.Engine((g['params'][:12], g['params'][12:19],
                         g['params'][19:])) as eng:
        o = eng.model_eval(g['confs'])
    # strict relative error against the live reference's values
    assert rel_err(o['lnpsi'], g['lnpsi']) < TOL
    assert rel_err(o['energy'], g['energy']) < TOL
    assert maxnorm_err(o['drift'], g['drift']) < TOL


@pytest.mark.parametrize('name,nconf', [('lat_n100', 3000), ('lat_n50', 5000),
                                        ('odd_n7', 4000), ('frac_n21', 2000),
                                        ('deep_n200', 300)])
def test_model_eval_vs_oracle_bulk(eng_mod, oracle, name, nconf):
    g = golden('model_' + name + '.npz')
    p = g['params']
    nop, size = int(p[3]), float(p[4])
    rng = np.random.default_rng(7)
    confs = np.zeros((nconf, 2, nop))
    confs[:, 0] = rng.random((nconf, nop)) * size
    ref = oracle.model_eval(p, confs)
    with eng_mod.Engine((p[:12], p[12:19], p[19:])) as eng:
        o = eng.model_eval(confs)
    for k in ('lnpsi', 'energy'):
        assert conditioned_rel_err(o[k], ref[k]) < TOL, k
        assert scaled_err(o[k], ref[k]) < TOL, k
    assert maxnorm_err(o['drift'], ref['drift']) < TOL


@pytest.mark.parametrize('name', golden_names('model_'))
def test_one_body_density_vs_reference(eng_mod, name):
    """g1(sz) of the live reference (qmc_base/jastrow/model.py:934-965) at
    offsets inside, at and far beyond the box."""
    g = golden(name)
    nobd = g['obd'].shape[0]
    with eng_mod.Engine((g['params'][:12], g['params'][12:19],
                         g['params'][19:])) as eng:
        o = eng.one_body_density(g['confs'][:nobd], g['obd_offsets'])
    assert o.shape == g['obd'].shape
    assert rel_err(o, g['obd']) < TOL


@pytest.mark.parametrize('name,nconf,noff', [
    ('lat_n100', 37, 50), ('lat_n50', 300, 7), ('odd_n7', 1000, 1),
    ('frac_n21', 11, 301), ('deep_n200', 9, 33), ('ideal_n8', 50, 16),
    ('strong_n10', 129, 128)])
def test_one_body_density_vs_oracle_bulk(eng_mod, oracle, name, nconf, noff):
    """Item counts that do not divide the CTA size: a CTA's items straddle
    several configurations, and the last CTA is ragged."""
    g = golden('model_' + name + '.npz')
    p = g['params']
    nop, size = int(p[3]), float(p[4])
    rng = np.random.default_rng(11)
    confs = np.zeros((nconf, 2, nop))
    confs[:, 0] = rng.random((nconf, nop)) * size
    offsets = (rng.random(noff) - 0.5) * 3 * size
    offsets[0] = 0.0
    ref = oracle.one_body_density(p, confs, offsets)
    with eng_mod.Engine((p[:12], p[12:19], p[19:])) as eng:
        o = eng.one_body_density(confs, offsets)
        assert eng.one_body_density(confs[:0], offsets).shape == (0, noff)
        assert eng.one_body_density(confs, offsets[:0]).shape == (nconf, 0)
    assert rel_err(o, ref) < TOL
    if not (p[10] != 0 and p[11] != 0):
        assert np.all(o[:, 0] == 1.0)       # zero displacement: exactly 1


@pytest.mark.parametrize('name', golden_names('model_'))
def test_physical_funcs_vs_reference(name):
    """The PhysicalFuncs mirror (gufunc signatures and broadcasting of
    qmc_base/jastrow/model.py:1007-1122) against the live reference."""
    from phd_qmclib_b200 import model
    g = golden(name)
    p = g['params']
    pf = model.PhysicalFuncs((p[:12], p[12:19], p[19:]))
    confs = g['confs']
    assert rel_err(pf.wf_abs_log(confs), g['lnpsi']) < TOL
    assert rel_err(pf.energy(confs), g['energy']) < TOL
    assert np.ndim(pf.energy(confs[0])) == 0
    nobd = g['obd'].shape[0]
    # (S,1) against (B,2,N) broadcasts to (S,B)
    obd = pf.one_body_density(g['obd_offsets'][:, None], confs[:nobd])
    assert obd.shape == (len(g['obd_offsets']), nobd)
    assert rel_err(obd.T, g['obd']) < TOL
    assert rel_err(pf.one_body_density(g['obd_offsets'][2], confs[1]),
                   g['obd'][1, 2]) < TOL
    fk = pf.fourier_density(g['kz_set'], confs.reshape((2, -1) + confs.shape[1:]))
    assert fk.shape == (2, confs.shape[0] // 2, len(g['kz_set']))
    assert np.max(np.abs(fk.reshape(g['fdk_k'].shape) - g['fdk_k'])) \
        < TOL * confs.shape[2]
    pf.engine.close()


@pytest.mark.parametrize('name', golden_names('cswf_'))
def test_cs_optimizer_vs_reference(name):
    """principal_function of the GPU optimiser against the reference's
    correlated-sampling variance at each frozen trial cutoff."""
    from phd_qmclib_b200 import model
    g = golden(name)
    kw = dict(zip([str(k) for k in g['spec_keys']], g['spec_vals']))
    spec = model.Spec(**kw)
    opt = model.CSWFOptimizer(spec, g['confs'], g['ini_lnpsi'])
    for k, rm in enumerate(g['cutoffs']):
        var = opt.principal_function(rm)
        scale = max(g['variance'][k], TOL * np.mean(g['energy'][k] ** 2))
        assert abs(var - g['variance'][k]) < 1e-9 * scale, (rm, var)
        ln, en = opt.wf_abs_log_and_energy_set(opt.update_spec(rm).cfc_spec)
        assert scaled_err(ln, g['lnpsi'][k]) < TOL
        assert scaled_err(en, g['energy'][k]) < TOL
    # the handle's own parameters are untouched by trial evaluations
    res = opt.engine.cs_variance(None, want_sets=True)
    assert scaled_err(res['wf_abs_log'], g['ini_lnpsi']) < TOL
    # ini_lnpsi = NULL: evaluated on the device with the current parameters
    opt.engine.cs_load(g['confs'], None)
    v2 = opt.engine.cs_variance(opt.update_spec(g['cutoffs'][1]))['variance']
    scale = max(g['variance'][1], TOL * np.mean(g['energy'][1] ** 2))
    assert abs(v2 - g['variance'][1]) < 1e-9 * scale
    # swapping the parameters of the live handle
    opt.engine.set_model_params(opt.update_spec(g['cutoffs'][-1]))
    o = opt.engine.model_eval(g['confs'], want=('lnpsi', 'energy'))
    assert scaled_err(o['lnpsi'], g['lnpsi'][-1]) < TOL
    assert scaled_err(o['energy'], g['energy'][-1]) < TOL
    opt.engine.close()


def test_cs_optimizer_exec_finds_minimum():
    """exec(): differential evolution on the host over the device objective
    (mrbp_qmc/model.py:929-942) lands on the minimum of a scan."""
    from phd_qmclib_b200 import model
    g = golden('cswf_odd_n7.npz')
    kw = dict(zip([str(k) for k in g['spec_keys']], g['spec_vals']))
    spec = model.Spec(**kw)
    opt = model.CSWFOptimizer(spec, g['confs'], g['ini_lnpsi'])
    lo, hi = opt.principal_function_bounds[0]
    scan = np.linspace(lo, hi, 60)
    vals = np.array([opt.principal_function(r) for r in scan])
    best = opt.exec(seed=1, maxiter=30, popsize=12, tol=1e-8)
    assert isinstance(best, model.Spec)
    assert lo <= best.tbf_contact_cutoff <= hi
    assert opt.principal_function(best.tbf_contact_cutoff) <= vals.min() + 1e-9
    opt.engine.close()


def _run_both(eng_mod, oracle, p, ini, wmax, target, dt, nwc, seed, nts,
              nblocks, energy_mode=0):
    size = float(p[4])
    st = oracle.DMCState(p, ini, wmax)
    eng = eng_mod.Engine((p[:12], p[12:19], p[19:]))
    dp = eng.dmc_params(dt, wmax, target, nwc, seed, 0.0, size,
                        energy_mode=energy_mode)
    eng.dmc_init(dp, ini)
    res = []
    for _ in range(nblocks):
        a = st.run_block(seed, dt, target, nwc, nts, 0.0, size,
                         energy_mode=energy_mode)
        b = eng.dmc_run_block(nts)
        res.append((a, b))
    return st, eng, res


@pytest.mark.parametrize('name,n_ini,wmax,nts', [
    ('ll_n16', 48, 64, 8), ('lat_n50', 40, 56, 6), ('odd_n7', 64, 96, 8),
    ('defects_n20', 32, 48, 6), ('lat_n100', 24, 32, 4),
    ('strong_n10', 40, 64, 6)])
@pytest.mark.parametrize('energy_mode', [0, 1])
def test_dmc_blocks_vs_oracle(eng_mod, oracle, name, n_ini, wmax, nts,
                              energy_mode):
    """Same seed, same Philox streams: branching decisions must be identical
    and every per-step series must agree to rounding."""
    g = golden('model_' + name + '.npz')
    p = g['params']
    nop, size = int(p[3]), float(p[4])
    rng = np.random.default_rng(3)
    ini = np.zeros((n_ini, 2, nop))
    ini[:, 0] = rng.random((n_ini, nop)) * size
    st, eng, res = _run_both(eng_mod, oracle, p, ini, wmax, n_ini, 1e-3,
                             0.125, 42, nts, 2, energy_mode)
    for a, b in res:
        assert np.array_equal(a['num_walkers'], b['num_walkers'])
        for k in ('energy', 'weight', 'ref_energy', 'accum_energy'):
            assert rel_err(b[k], a[k]) < 1e-10, k
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
    eng.close()


def test_dmc_capacity_hit(eng_mod, oracle):
    """Capacity truncation (quirk Q8) follows the reference's order."""
    g = golden('model_ll_n16.npz')
    p = g['params']
    rng = np.random.default_rng(5)
    ini = np.zeros((30, 2, 16))
    ini[:, 0] = rng.random((30, 16)) * 16
    # a far-too-low reference energy makes the population explode
    st = oracle.DMCState(p, ini, 40, ref_energy=400.0)
    eng = eng_mod.Engine((p[:12], p[12:19], p[19:]))
    dp = eng.dmc_params(5e-3, 40, 30, 0.0, 9, 0.0, 16.0)
    eng.dmc_init(dp, ini, ref_energy=400.0)
    a = st.run_block(9, 5e-3, 30, 0.0, 4, 0.0, 16.0)
    b = eng.dmc_run_block(4)
    assert a['num_walkers'].max() == 40
    assert np.array_equal(a['num_walkers'], b['num_walkers'])
    assert rel_err(b['energy'], a['energy']) < 1e-10
    assert eng.dmc_scalars().capacity_hits > 0
    eng.close()


@pytest.mark.parametrize('name,nconf,modes', [('lat_n50', 300, 50),
                                              ('odd_n7', 500, 37),
                                              ('deep_n200', 40, 400)])
def test_fourier_density_vs_oracle(eng_mod, oracle, name, nconf, modes):
    g = golden('model_' + name + '.npz')
    p = g['params']
    nop, size = int(p[3]), float(p[4])
    rng = np.random.default_rng(11)
    confs = np.zeros((nconf, 2, nop))
    confs[:, 0] = rng.random((nconf, nop)) * size
    ref = oracle.fourier_density(p, confs, modes)
    with eng_mod.Engine((p[:12], p[12:19], p[19:])) as eng:
        got = eng.fourier_density(confs, modes)
    # k = 0: |rho|^2 = N^2, Re = N exactly (SURVEY 8c sanity value)
    assert np.allclose(got[:, 0, 0], nop ** 2, rtol=1e-14)
    assert np.allclose(got[:, 0, 1], nop, rtol=1e-14)
    assert np.max(np.abs(got - ref)) < 1e-12 * nop ** 2


@pytest.mark.parametrize('name', golden_names('model_'))
def test_fourier_density_vs_reference(eng_mod, name):
    g = golden(name)
    p = g['params']
    nop = int(p[3])
    with eng_mod.Engine((p[:12], p[12:19], p[19:])) as eng:
        got = eng.fourier_density(g['confs'], int(g['num_modes']))
    assert np.max(np.abs(got - g['ssf'])) < 1e-12 * nop ** 2


@pytest.mark.parametrize('pure', [True, False])
@pytest.mark.parametrize('name,n_ini,wmax,nts,modes,bins', [
    ('ll_n16', 48, 64, 8, 16, 64), ('lat_n50', 40, 56, 6, 50, 400),
    ('odd_n7', 64, 96, 9, 21, 33)])
def test_dmc_estimators_vs_oracle(eng_mod, oracle, name, n_ini, wmax, nts,
                                  modes, bins, pure):
    """S(k) and density series of two blocks (forward walking restarts at
    every block; the window pfw = nts - 2 also exercises plain transport)."""
    g = golden('model_' + name + '.npz')
    p = g['params']
    nop, size = int(p[3]), float(p[4])
    rng = np.random.default_rng(4)
    ini = np.zeros((n_ini, 2, nop))
    ini[:, 0] = rng.random((n_ini, nop)) * size
    pfw = nts - 2
    st = oracle.DMCState(p, ini, wmax)
    ssf = dict(num=modes, pure=pure, pfw=pfw, iter=np.zeros((nts, modes, 3)),
               aux=np.zeros((2, wmax, modes, 3)))
    den = dict(num=bins, pure=pure, pfw=pfw, iter=np.zeros((nts, bins)),
               aux=np.zeros((2, wmax, bins)))
    eng = eng_mod.Engine((p[:12], p[12:19], p[19:]))
    dp = eng.dmc_params(2e-3, wmax, n_ini, 0.25, 5, 0.0, size,
                        ssf=(modes, pure, pfw), density=(bins, pure, pfw))
    eng.dmc_init(dp, ini)
    for blk in range(3):
        est = blk > 0                      # block 0 plays the burn-in role
        for d in (ssf, den):
            d['iter'][:] = 0
            d['aux'][:] = 0
        a = st.run_block(5, 2e-3, n_ini, 0.25, nts, 0.0, size, eval_est=est,
                         ssf=ssf, density=den)
        e_den, e_ssf = np.zeros((nts, bins)), np.zeros((nts, modes, 3))
        b = eng.dmc_run_block(nts, eval_estimators=est, density=e_den,
                              ssf=e_ssf)
        assert np.array_equal(a['num_walkers'], b['num_walkers'])
        if est:
            assert np.allclose(e_den, den['iter'], rtol=1e-13, atol=1e-13)
            assert np.max(np.abs(e_ssf - ssf['iter'])) \
                < 1e-11 * np.max(np.abs(ssf['iter']))
    eng.close()


@pytest.mark.parametrize('name,nch,ns,modes,spread,proposal', [
    ('ll_n16', 40, 12, 16, 0.5, 0), ('lat_n50', 23, 8, 50, 0.125, 0),
    ('odd_n7', 64, 16, 11, 0.8, 0), ('defects_n20', 30, 10, 0, 0.2, 0),
    ('ideal_n8', 16, 10, 8, 0.3, 0), ('lat_n50', 23, 8, 10, 0.05, 1),
    ('frac_n21', 31, 9, 5, 0.1, 1)])
def test_vmc_blocks_vs_oracle(eng_mod, oracle, name, nch, ns, modes, spread,
                              proposal):
    """Same Philox streams: the accept/reject sequence must be identical and
    every series must agree to rounding, over three consecutive blocks."""
    g = golden('model_' + name + '.npz')
    p = g['params']
    nop, size = int(p[3]), float(p[4])
    rng = np.random.default_rng(21)
    ini = np.zeros((nch, 2, nop))
    ini[:, 0] = rng.random((nch, nop)) * size
    cur = ini.copy()
    ln = oracle.model_eval(p, cur, want=('lnpsi',))['lnpsi']
    eprev = np.zeros(nch)
    sprev = np.zeros((nch, max(modes, 1), 3))
    eng = eng_mod.Engine((p[:12], p[12:19], p[19:]))
    eng.vmc_init(ini, spread, 77, 0.0, size, ssf_num_modes=modes,
                 chain_offset=5, proposal=proposal)
    step0 = 0
    for b in range(3):
        first = b == 0
        a = oracle.vmc_block(p, 77, spread, 0.0, size, cur, ln, eprev,
                             sprev if modes else None, modes, ns, step0,
                             first, chain_offset=5, proposal=proposal)
        step0 += ns - (1 if first else 0)
        o = eng.vmc_run_block(ns, series=True, sums=True)
        assert np.array_equal(o['move_stat'], a['stat'])
        # ln|Psi| and E_L cross zero in these small systems: strict relative
        # error away from the zero crossings, scaled error everywhere
        assert conditioned_rel_err(o['lnpsi'], a['lnpsi']) < 1e-11
        assert scaled_err(o['lnpsi'], a['lnpsi']) < 1e-12
        assert scaled_err(o['energy'], a['energy']) < 1e-11
        assert np.allclose(o['accept_rate'], a['accept_rate'], rtol=0,
                           atol=1e-15)
        assert np.allclose(o['sum_energy'][:, 0], a['energy'].sum(axis=1),
                           rtol=1e-11)
        if modes:
            assert np.max(np.abs(o['ssf'] - a['ssf'])) < 1e-9
            assert np.allclose(o['sum_ssf'], a['ssf'].sum(axis=1),
                               rtol=1e-9, atol=1e-8)
    confs, lnpsi = eng.vmc_get_state()
    assert np.allclose(confs[:, 0], cur[:, 0], rtol=0, atol=1e-12)
    assert rel_err(lnpsi, ln) < 1e-11
    eng.close()


def test_vmc_reinit_reuses_buffers_and_fills_caller_arrays(eng_mod, oracle):
    """`vmc_init` of the same shape on a live engine (the end-to-end loop of
    bench.py) must behave like a fresh engine; results can land in
    caller-owned page-locked arrays; a start outside [0, L] (legal: the
    reference only wraps proposals) turns the node tables off, not the
    answer."""
    g = golden('model_lat_n50.npz')
    p = g['params']
    nop, size = int(p[3]), float(p[4])
    nch, ns, M = 37, 6, 9
    rng = np.random.default_rng(5)
    eng = eng_mod.Engine((p[:12], p[12:19], p[19:]))
    out = {'accept_rate': eng_mod.pinned_empty((nch,)),
           'sum_energy': eng_mod.pinned_empty((nch, 2)),
           'sum_ssf': eng_mod.pinned_empty((nch, M, 3)),
           'lnpsi': eng_mod.pinned_empty((nch, ns)),
           'energy': eng_mod.pinned_empty((nch, ns)),
           'move_stat': eng_mod.pinned_empty((nch, ns), np.uint8),
           'ssf': eng_mod.pinned_empty((nch, ns, M, 3))}
    state = (eng_mod.pinned_empty((nch, 2, nop)), eng_mod.pinned_empty((nch,)))
    for trial, shift in enumerate((0.0, 0.0, 0.75)):
        ini = np.zeros((nch, 2, nop))
        ini[:, 0] = rng.random((nch, nop)) * size
        ini[3, 0, 7] = size + shift if shift else ini[3, 0, 7]
        ini[5, 0, 0] = -shift if shift else ini[5, 0, 0]
        eng.vmc_init(ini, 0.2, 31 + trial, 0.0, size, ssf_num_modes=M,
                     chain_offset=2)
        ln = oracle.model_eval(p, ini, want=('lnpsi',))['lnpsi']
        cur = ini.copy()
        a = oracle.vmc_block(p, 31 + trial, 0.2, 0.0, size, cur, ln,
                             np.zeros(nch), np.zeros((nch, M, 3)), M, ns, 0,
                             True, chain_offset=2)
        o = eng.vmc_run_block(ns, series=True, sums=True, out=out)
        for key in out:
            assert o[key] is out[key]
        assert np.array_equal(o['move_stat'], a['stat'])
        assert scaled_err(o['lnpsi'], a['lnpsi']) < 1e-12
        assert scaled_err(o['energy'], a['energy']) < 1e-11
        assert np.max(np.abs(o['ssf'] - a['ssf'])) < 1e-9
        assert np.allclose(o['sum_ssf'], a['ssf'].sum(axis=1), rtol=1e-9,
                           atol=1e-8)
        confs, lnpsi = eng.vmc_get_state(out=state)
        assert confs is state[0] and lnpsi is state[1]
        assert np.allclose(confs[:, 0], cur[:, 0], rtol=0, atol=1e-12)
    with pytest.raises(ValueError):
        eng.vmc_run_block(ns, series=False, sums=True,
                          out={'sum_energy': np.empty((nch, 3))})
    with pytest.raises(ValueError):
        eng.vmc_get_state(out=(np.empty((nch, 2, nop + 1)), np.empty(nch)))
    eng.close()


@pytest.mark.parametrize('kwargs', [
    # every cell is a defect (defects_sep == 1 with V_defect != V0)
    dict(lattice_depth=40.0, lattice_ratio=1, interaction_strength=2,
         boson_number=10, supercell_size=10, tbf_contact_cutoff=2.5,
         num_defects=10, defect_magnitude=15.0),
    # r_m = L/2: every pair is on the short-range branch
    dict(lattice_depth=20.0, lattice_ratio=0.5, interaction_strength=3,
         boson_number=12, supercell_size=9, tbf_contact_cutoff=4.5),
    # very weak interaction: k2 tiny, large unit-conversion factors
    dict(lattice_depth=10.0, lattice_ratio=1, interaction_strength=1e-6,
         boson_number=9, supercell_size=9, tbf_contact_cutoff=2.0),
    # free gas (no lattice) and a single particle block
    dict(lattice_depth=0.0, lattice_ratio=1, interaction_strength=5,
         boson_number=3, supercell_size=4, tbf_contact_cutoff=1.0),
])
def test_model_eval_edge_specs_vs_oracle(eng_mod, oracle, kwargs):
    from phd_qmclib_b200 import model
    spec = model.Spec(**kwargs)
    p = model.param_block(spec)
    nop, size = spec.boson_number, spec.supercell_size
    rng = np.random.default_rng(17)
    confs = np.zeros((400, 2, nop))
    confs[:, 0] = rng.random((400, nop)) * size
    if spec.tbf_contact_cutoff < 0.5 * size:
        # (with r_m = L/2 a pair at exactly L/2 sits on the branch cut of a
        # trial function whose far branch is empty -- DESIGN.md, deviation D3)
        confs[0, 0] = np.linspace(0, size, nop, endpoint=False)
    ref = oracle.model_eval(p, confs)
    with eng_mod.Engine(spec) as eng:
        o = eng.model_eval(confs)
    assert scaled_err(o['lnpsi'], ref['lnpsi']) < TOL
    assert scaled_err(o['energy'], ref['energy']) < TOL
    assert maxnorm_err(o['drift'], ref['drift']) < TOL


def test_dmc_restart_is_bit_exact(eng_mod):
    """qmcb_dmc_get_next -> qmcb_dmc_set_state on a fresh handle continues the
    run bit for bit (what a checkpoint/resume of the evolved population
    needs: walkers, weights, the per-slot stale energies of quirk Q1, the
    running totals and the step counter that keys the RNG)."""
    g = golden('model_lat_n50.npz')
    p = g['params']
    spec = (p[:12], p[12:19], p[19:])
    rng = np.random.default_rng(8)
    ini = np.zeros((60, 2, 50))
    ini[:, 0] = rng.random((60, 50)) * 50
    a = eng_mod.Engine(spec)
    dp = a.dmc_params(1e-3, 96, 60, 0.25, 13, 0.0, 50.0)
    a.dmc_init(dp, ini)
    a.dmc_run_block(7)
    nx = a.dmc_get_next()
    want = a.dmc_run_block(9)
    b = eng_mod.Engine(spec)
    b.dmc_set_state(dp, nx['confs'], nx['energy'], nx['weight'],
                    nx['scalars'], slot_energy=nx['slot_energy'])
    got = b.dmc_run_block(9)
    for k in want:
        assert np.array_equal(want[k], got[k]), k
    sa, sb = a.dmc_get_state(), b.dmc_get_state()
    for k in ('confs', 'energy', 'weight', 'mask', 'cloning_ref'):
        assert np.array_equal(sa[k], sb[k]), k
    a.close()
    b.close()


@pytest.mark.parametrize('nop', [1, 2, 5, 33, 257, 400, 1000])
def test_size_extremes_vs_oracle(eng_mod, oracle, nop):
    """From a single particle (no pairs at all) to a walker that fills a
    whole 256-thread CTA (nb = 250 particle blocks): lnPsi, E_L, drift, rho_k,
    g1 and two DMC steps against the oracle."""
    from phd_qmclib_b200 import model
    size = float(max(nop, 2))
    spec = model.Spec(3 * np.pi ** 2, 1, 2.5, nop, size, 0.25 * size)
    p = model.param_block(spec)
    rng = np.random.default_rng(nop)
    nconf = 6 if nop >= 257 else 40
    confs = np.zeros((nconf, 2, nop))
    confs[:, 0] = rng.random((nconf, nop)) * size
    ref = oracle.model_eval(p, confs)
    with eng_mod.Engine(spec) as eng:
        o = eng.model_eval(confs)
        assert scaled_err(o['lnpsi'], ref['lnpsi']) < TOL
        assert scaled_err(o['energy'], ref['energy']) < TOL
        assert maxnorm_err(o['drift'], ref['drift']) < TOL
        # empty batches are a no-op, not an error
        e = eng.model_eval(confs[:0])
        assert e['lnpsi'].shape == (0,) and e['drift'].shape == (0, nop)
        modes = 5
        s = eng.fourier_density(confs, modes)
        assert np.max(np.abs(s - oracle.fourier_density(p, confs, modes))) \
            < 1e-12 * nop ** 2
        offs = np.array([0.0, 0.3, -1.2 * size])
        g1 = eng.one_body_density(confs[:3], offs)
        assert rel_err(g1, oracle.one_body_density(p, confs[:3], offs)) < 1e-11
        if nop == 1000:
            # one configuration per CTA, several CTAs per configuration
            offs = np.linspace(-0.5 * size, 0.5 * size, 130)
            g1 = eng.one_body_density(confs[:2], offs)
            assert rel_err(g1, oracle.one_body_density(p, confs[:2], offs)) \
                < 1e-11
        wmax = nconf + 4
        dp = eng.dmc_params(1e-3, wmax, nconf, 0.125, 5, 0.0, size)
        eng.dmc_init(dp, confs)
        b = eng.dmc_run_block(2)
    st = oracle.DMCState(p, confs, wmax)
    a = st.run_block(5, 1e-3, nconf, 0.125, 2, 0.0, size)
    assert np.array_equal(a['num_walkers'], b['num_walkers'])
    for k in ('energy', 'weight', 'ref_energy', 'accum_energy'):
        assert rel_err(b[k], a[k]) < 1e-9, k


def test_too_many_particles_is_an_error(eng_mod):
    """Beyond 256 particle blocks a walker no longer fits one CTA: the
    engine says so at creation instead of computing something else."""
    from phd_qmclib_b200 import model
    from phd_qmclib_b200._lib import EngineError
    spec = model.Spec(10.0, 1, 2, 1028, 1028.0, 257.0)
    with pytest.raises(EngineError, match='too large'):
        eng_mod.Engine(spec)


def test_block_graph_equals_stream_launches(eng_mod):
    """A block without estimators replays a captured CUDA graph; with
    per-launch events (set_profiling) the same block is launched kernel by
    kernel.  Both must give the same bits, across a change of the block
    length (graph rebuilt) and a restart (graph kept or rebuilt)."""
    g = golden('model_lat_n50.npz')
    p = g['params']
    spec = (p[:12], p[12:19], p[19:])
    rng = np.random.default_rng(31)
    ini = np.zeros((70, 2, 50))
    ini[:, 0] = rng.random((70, 50)) * 50
    a, b = eng_mod.Engine(spec), eng_mod.Engine(spec)
    dp = a.dmc_params(1e-3, 128, 70, 0.25, 21, 0.0, 50.0)
    a.dmc_init(dp, ini)
    b.dmc_init(dp, ini)
    b.set_profiling(True)
    for nts in (5, 5, 7, 1, 7):
        x, y = a.dmc_run_block(nts), b.dmc_run_block(nts)
        for k in x:
            assert np.array_equal(x[k], y[k]), (nts, k)
        assert a.last_block_stats()['launches'] == 1 + 3 * nts
    nx = a.dmc_get_next()
    dp2 = a.dmc_params(2e-3, 128, 70, 0.25, 21, 0.0, 50.0)   # new time step
    for eng in (a, b):
        eng.dmc_set_state(dp2, nx['confs'], nx['energy'], nx['weight'],
                          nx['scalars'], slot_energy=nx['slot_energy'])
    x, y = a.dmc_run_block(7), b.dmc_run_block(7)
    for k in x:
        assert np.array_equal(x[k], y[k]), k
    sa, sb = a.dmc_get_state(), b.dmc_get_state()
    for k in ('confs', 'energy', 'weight', 'mask', 'cloning_ref'):
        assert np.array_equal(sa[k], sb[k]), k
    a.close()
    b.close()


@pytest.mark.parametrize('pure', [True, False])
def test_density_bin_edges_vs_oracle(eng_mod, oracle, pure):
    """Positions exactly on bin edges and one ulp to either side, with a bin
    width that is not a binary fraction: the engine's floor division (an fma
    residual instead of the reference's fmod-based Python `//`) must land in
    the same bin as the oracle, which restates the reference op for op."""
    g = golden('model_frac_n21.npz')
    p = g['params']
    nop, size = int(p[3]), float(p[4])
    bins, n_ini, wmax, nts = 37, 40, 64, 3
    width = size / bins
    rng = np.random.default_rng(12)
    # distinct bins within a walker: two particles at exactly the same
    # position have no defined sign(d) (the reference gives both +f2'/f2, an
    # antisymmetric pair evaluation gives +-), which is not what this test
    # is about
    k = np.argsort(rng.random((n_ini, bins)), axis=1)[:, :nop] \
        .astype(np.float64)
    z = k * width
    z[1::3] = np.nextafter(z[1::3], np.inf)
    z[2::3] = np.nextafter(z[2::3], -np.inf)
    z = np.clip(z, 0.0, np.nextafter(size, 0.0))
    ini = np.zeros((n_ini, 2, nop))
    ini[:, 0] = z
    st = oracle.DMCState(p, ini, wmax)
    den = dict(num=bins, pure=pure, pfw=nts, iter=np.zeros((nts, bins)),
               aux=np.zeros((2, wmax, bins)))
    eng = eng_mod.Engine((p[:12], p[12:19], p[19:]))
    dp = eng.dmc_params(1e-4, wmax, n_ini, 0.25, 5, 0.0, size,
                        density=(bins, pure, nts))
    eng.dmc_init(dp, ini)
    a = st.run_block(5, 1e-4, n_ini, 0.25, nts, 0.0, size, eval_est=True,
                     density=den)
    e_den = np.zeros((nts, bins))
    b = eng.dmc_run_block(nts, eval_estimators=True, density=e_den)
    assert np.array_equal(a['num_walkers'], b['num_walkers'])
    # step 0 histograms the initial positions themselves: integer counts
    assert np.array_equal(e_den[0], den['iter'][0])
    assert np.allclose(e_den, den['iter'], rtol=1e-13, atol=1e-13)
    eng.close()


def test_on_device_reblocking_vs_oracle(eng_mod, oracle):
    """qmcb_dmc_reblock_*: the accumulated on-the-fly reblocking tables of
    the five per-step series equal the reference's on_the_fly_obj_create of
    each shipped block series, accumulated with on_the_fly_obj_data_update
    -- bit for bit; block lengths 12 and 16 (orders 3 and 4)."""
    g = golden('model_ll_n16.npz')
    p = g['params']
    rng = np.random.default_rng(2)
    ini = np.zeros((200, 2, 16))
    ini[:, 0] = rng.random((200, 16)) * 16
    eng = eng_mod.Engine((p[:12], p[12:19], p[19:]))
    dp = eng.dmc_params(2e-3, 320, 200, 0.125, 3, 0.0, 16.0)
    eng.dmc_init(dp, ini)
    eng.dmc_run_block(8)                    # before the reset: not counted
    eng.dmc_reblock_reset(4)
    want = {}
    for nts in (12, 16, 12, 1):
        blk = eng.dmc_run_block(nts)
        for name in eng.REBLOCK_SERIES:
            t_ = oracle.otf_create(blk[name].astype(np.float64))
            full = np.zeros(5, dtype=oracle.OTF_DTYPE)
            full['BLOCK_SIZE'] = 1 << np.arange(5)
            k = min(len(t_), 5)
            for f in ('MEANS', 'MEANS_SQR', 'NUM_BLOCKS'):
                full[f][:k] = t_[f][:k]
            if name in want:
                oracle.otf_update(want[name], full)
            else:
                want[name] = full
    got = eng.dmc_reblock_get()
    for name in eng.REBLOCK_SERIES:
        assert got[name].dtype == oracle.OTF_DTYPE
        for f in oracle.OTF_DTYPE.names:
            assert np.array_equal(got[name][f], want[name][f]), (name, f)
    assert got['energy']['NUM_BLOCKS'].tolist() == [41, 20, 10, 4, 1]
    eng.dmc_reblock_reset(None)             # switched off again
    eng.dmc_run_block(4)
    with pytest.raises(Exception):
        eng.dmc_reblock_get()
    eng.close()


@pytest.mark.parametrize('name', ['lat_n50', 'defects_n20', 'odd_n7'])
def test_shifted_recast_interval_takes_the_exact_path(eng_mod, oracle, name):
    """The node tables of the per-particle transcendentals cover [0, L]; a
    sampling that recasts into another interval (here [-L/3, 2L/3]) runs the
    exact sincospi / exp variants of the DMC step and VMC block kernels.
    Same parity bar against the oracle."""
    g = golden('model_' + name + '.npz')
    p = g['params']
    nop, size = int(p[3]), float(p[4])
    lo, hi = -size / 3, 2 * size / 3
    rng = np.random.default_rng(41)
    n_ini, wmax, nts = 48, 80, 6
    ini = np.zeros((n_ini, 2, nop))
    ini[:, 0] = lo + rng.random((n_ini, nop)) * size
    st = oracle.DMCState(p, ini, wmax)
    eng = eng_mod.Engine((p[:12], p[12:19], p[19:]))
    dp = eng.dmc_params(1e-3, wmax, n_ini, 0.25, 6, lo, hi)
    eng.dmc_init(dp, ini)
    for _ in range(2):
        a = st.run_block(6, 1e-3, n_ini, 0.25, nts, lo, hi)
        b = eng.dmc_run_block(nts)
        assert np.array_equal(a['num_walkers'], b['num_walkers'])
        for k in ('energy', 'weight', 'ref_energy', 'accum_energy'):
            assert rel_err(b[k], a[k]) < 1e-10, k
    nw = st.num_walkers
    nx = eng.dmc_get_next()
    assert np.allclose(nx['confs'][:, 0], st.prev['confs'][:nw, 0], rtol=0,
                       atol=1e-9)
    assert nx['confs'][:, 0].min() >= lo and nx['confs'][:, 0].max() <= hi
    # VMC on the same interval
    nch, ns, modes = 24, 10, 6
    cur = ini[:nch].copy()
    ln = oracle.model_eval(p, cur, want=('lnpsi',))['lnpsi']
    eprev, sprev = np.zeros(nch), np.zeros((nch, modes, 3))
    eng.vmc_init(cur, 0.2, 8, lo, hi, ssf_num_modes=modes)
    step0 = 0
    for blk in range(2):
        first = blk == 0
        a = oracle.vmc_block(p, 8, 0.2, lo, hi, cur, ln, eprev, sprev, modes,
                             ns, step0, first)
        step0 += ns - (1 if first else 0)
        o = eng.vmc_run_block(ns, series=True)
        assert np.array_equal(o['move_stat'], a['stat'])
        assert scaled_err(o['lnpsi'], a['lnpsi']) < 1e-11
        assert scaled_err(o['energy'], a['energy']) < 1e-11
        assert np.max(np.abs(o['ssf'] - a['ssf'])) < 1e-9
    eng.close()
