"""CPU: the C oracle against outputs of the LIVE reference (tests/golden/,
made by oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

from conftest import golden, golden_names, maxnorm_err, rel_err

TOL = 1e-13     # the oracle follows the reference op for op: expect ~0


@pytest.mark.parametrize('name', golden_names('model_'))
def test_model_eval(oracle, name):
    g = golden(name)
    o = oracle.model_eval(g['params'], g['confs'])
    assert rel_err(o['lnpsi'], g['lnpsi']) < TOL
    assert rel_err(o['energy'], g['energy']) < TOL
    assert maxnorm_err(o['drift'], g['drift']) < TOL
    s = oracle.fourier_density(g['params'], g['confs'], int(g['num_modes']))
    assert np.max(np.abs(s - g['ssf'])) < 1e-12 * g['confs'].shape[2] ** 2
    nobd = g['obd'].shape[0]
    obd = oracle.one_body_density(g['params'], g['confs'][:nobd],
                                  g['obd_offsets'])
    assert rel_err(obd, g['obd']) < 1e-12
    fk = oracle.fourier_density_k(g['params'], g['confs'], g['kz_set'])
    assert np.max(np.abs(fk - g['fdk_k'])) < 1e-12 * g['confs'].shape[2]


@pytest.mark.parametrize('name', golden_names('cswf_'))
def test_cs_variance(oracle, name):
    """Correlated-sampling objective: ln|Psi|, E_L of the fixed set under
    each trial cutoff and the reference's weighed_variance."""
    g = golden(name)
    for k, block in enumerate(g['trial_params']):
        o = oracle.model_eval(block, g['confs'])
        assert rel_err(o['lnpsi'], g['lnpsi'][k]) < TOL
        assert rel_err(o['energy'], g['energy'][k]) < TOL
        var, _ = oracle.weighed_variance(o['lnpsi'], g['ini_lnpsi'],
                                         o['energy'])
        scale = max(g['variance'][k], 1e-12 * np.mean(g['energy'][k] ** 2))
        assert abs(var - g['variance'][k]) < 1e-9 * scale


@pytest.mark.parametrize('name', golden_names('dmc_step_'))
def test_evolve_state(oracle, name):
    """One reference evolve_state call: gather by parent, recast, E'/F',
    stale-slot weight (Q1), cloning into `actual`, dead-slot mask."""
    g = golden(name)
    wmax = int(g['max_num_walkers'])
    st = oracle.DMCState(g['params'], g['ini_confs'], wmax)
    # build_state parity (mrbp_qmc/dmc.py:268-328)
    assert rel_err(st.prev['confs'], g['ini_state_confs']) < TOL or \
        maxnorm_err(st.prev['confs'].reshape(wmax, -1),
                    g['ini_state_confs'].reshape(wmax, -1)) < TOL
    assert np.array_equal(st.prev['mask'].astype(bool), g['ini_state_mask'])
    assert rel_err(st.prev['energy'], g['ini_state_energy']) < TOL
    assert abs(st.scal[0] - float(g['ini_ref_energy'])) \
        < TOL * abs(float(g['ini_ref_energy']))
    st.act['energy'][:] = g['act_energy_in']
    st.ref[:] = g['cloning_ref']
    nop = g['ini_confs'].shape[2]
    st.evolve(int(g['num_walkers']), float(g['time_step']),
              float(g['ref_energy']), np.zeros((wmax, nop)),
              float(g['z_min']), float(g['z_max']))
    nw = int(g['num_walkers'])
    assert np.array_equal(st.act['mask'].astype(bool), g['act_mask'])
    assert np.array_equal(st.act['confs'][:nw], g['act_confs'][:nw])
    assert np.array_equal(st.act['energy'], g['act_energy'])
    assert np.array_equal(st.act['weight'][:nw], g['act_weight'][:nw])
    assert np.array_equal(st.next['confs'][:nw, 0], g['next_confs'][:nw, 0])
    assert maxnorm_err(st.next['confs'][:nw, 1], g['next_confs'][:nw, 1]) < TOL
    assert rel_err(st.next['energy'][:nw], g['next_energy'][:nw]) < TOL
    assert rel_err(st.next['weight'][:nw], g['next_weight'][:nw]) < TOL
    # Q1: two children of one parent carry different weights.
    assert g['cloning_ref'][0] == g['cloning_ref'][1]
    assert st.next['weight'][0] != st.next['weight'][1]


@pytest.mark.parametrize('tag', ['plain', 'capped', 'dying'])
def test_branch(oracle, tag):
    g = golden('dmc_branch.npz')
    w, u = g[f'{tag}_weights'], g[f'{tag}_uniforms']
    wmax = int(g[f'{tag}_wmax'])
    nw, ref = oracle.branch(w, len(w), wmax, u)
    assert nw == int(g[f'{tag}_num_walkers'])
    assert np.array_equal(ref[:nw], g[f'{tag}_cloning_ref'][:nw])
    if tag == 'capped':
        assert nw == wmax


@pytest.mark.parametrize('name', golden_names('dmc_blocks_'))
def test_dmc_blocks(oracle, name):
    """The reference's blocks() (serial, seeded) replayed with the exact
    random numbers it consumed: E_ref recurrence, buffer rotation, Q1
    dynamics, S(k) and density estimators over several steps."""
    g = golden(name)
    wmax, nts, nb = int(g['max_num_walkers']), int(g['nts']), int(g['nblocks'])
    nm, nbins, pure = int(g['num_modes']), int(g['num_bins']), bool(g['pure'])
    st = oracle.DMCState(g['params'], g['ini_confs'], wmax)
    ssf = dict(num=nm, pure=pure, pfw=nts, iter=np.zeros((nts, nm, 3)),
               aux=np.zeros((2, wmax, nm, 3)))
    den = dict(num=nbins, pure=pure, pfw=nts, iter=np.zeros((nts, nbins)),
               aux=np.zeros((2, wmax, nbins)))
    for b in range(nb):
        # qmc_base/dmc.py:901-909: per-block reset
        for d in (ssf, den):
            d['iter'][:] = 0
            d['aux'][:] = 0
        it = st.run_block(0, float(g['time_step']),
                          int(g['target_num_walkers']),
                          float(g['nwc_factor']), nts, float(g['z_min']),
                          float(g['z_max']), eval_est=True, ssf=ssf,
                          density=den,
                          uniforms_ext=g['uniforms'][b * nts:(b + 1) * nts],
                          normals_ext=g['normals'][b * nts:(b + 1) * nts])
        assert np.array_equal(it['num_walkers'], g['it_num_walkers'][b])
        for k in ('energy', 'weight', 'ref_energy', 'accum_energy'):
            assert rel_err(it[k], g['it_' + k][b]) < 1e-11, k
        assert np.max(np.abs(ssf['iter'] - g['it_ssf'][b])) \
            < 1e-10 * np.max(np.abs(g['it_ssf'][b]))
        assert np.allclose(den['iter'], g['it_density'][b, :, :, 0],
                           rtol=1e-12, atol=0)
    nw = int(g['last_num_walkers'])
    assert np.allclose(st.act['confs'][:nw], g['last_confs'][:nw],
                       rtol=1e-10, atol=1e-10)
    assert np.array_equal(st.ref[:nw], g['last_cloning_ref'][:nw])
    assert np.array_equal(st.act['mask'].astype(bool), g['last_mask'])


@pytest.mark.parametrize('name', golden_names('vmc_blocks_'))
def test_vmc_blocks(oracle, name):
    g = golden(name)
    ns, nb, nm = int(g['ns']), int(g['nblocks']), int(g['num_modes'])
    nop = g['ini_conf'].shape[1]
    cur = g['ini_conf'][None].copy()
    ln = oracle.model_eval(g['params'], cur, want=('lnpsi',))['lnpsi']
    assert rel_err(ln, g['ini_lnpsi']) < TOL
    eprev = np.zeros(1)
    sprev = np.zeros((1, nm, 3))
    uni = g['uniforms']
    for b in range(nb):
        first = b == 0
        lo = 0 if first else b * ns - 1
        ue = uni[lo:lo + ns - (1 if first else 0)]
        out = oracle.vmc_block(g['params'], 0, float(g['move_spread']),
                               float(g['z_min']), float(g['z_max']), cur, ln,
                               eprev, sprev, nm, ns, 0, first,
                               uniforms_ext=ue[:, None, :],
                               proposal=int(g['proposal'])
                               if 'proposal' in g.files else 0)
        assert np.array_equal(out['stat'][0].astype(bool), g['it_stat'][b])
        assert rel_err(out['lnpsi'][0], g['it_lnpsi'][b]) < 1e-12
        assert rel_err(out['energy'][0], g['it_energy'][b]) < 1e-11
        assert np.max(np.abs(out['ssf'][0] - g['it_ssf'][b])) < 1e-9
        assert out['accept_rate'][0] == pytest.approx(
            float(g['it_accept_rate'][b]), abs=1e-15)
    assert np.allclose(cur[0, 0], g['last_conf'][0], rtol=0, atol=1e-12)
    assert nop == int(g['params'][3])


def test_otf_reblocking(oracle):
    """The restatement of stats/reblock.py's on-the-fly reblocking against
    tables made by the live reference: bit for bit (same operations in the
    same order), for lengths that are and are not powers of two."""
    g = golden('reblock_otf.npz')
    for tag in 'abcd':
        otf = oracle.otf_create(g[f'{tag}_series'])
        for f in otf.dtype.names:
            assert np.array_equal(otf[f], g[f'{tag}_{f}']), (tag, f)
    blocks = g['acc_blocks']
    acc = oracle.otf_create(blocks[0])
    for b in blocks[1:]:
        oracle.otf_update(acc, oracle.otf_create(b))
    for f in acc.dtype.names:
        assert np.array_equal(acc[f], g[f'acc_{f}']), f
    # what the reference derives from such a table (OTFObject.mean / errors)
    nb = acc['NUM_BLOCKS'][acc['NUM_BLOCKS'] >= 2]
    k = len(nb)
    means = acc['MEANS'][:k] / nb
    var = nb * (acc['MEANS_SQR'][:k] / nb - means ** 2) / (nb - 1)
    assert means[0] == g['acc_mean']
    assert np.allclose(np.sqrt(var / nb), g['acc_errors'], rtol=1e-13)
