"""CPU: host-side logic -- model spec parameter derivation against the live
reference's values (frozen in tests/golden), and the sampler mirrors'
parameter/validation surface.  No GPU calls."""
import math

import numpy as np
import pytest

from conftest import golden
from specs import SPECS


@pytest.mark.parametrize('name', sorted(SPECS))
def test_spec_params_match_reference(name):
    """kp_ground_state_energy / tbf matching (reference mrbp_qmc/model.py:
    276-393, ideal.py:8-85) reproduce the reference's 25 scalars."""
    from phd_qmclib_b200 import model
    spec = model.Spec(**SPECS[name])
    got = model.param_block(spec)
    want = golden(f'model_{name}.npz')['params']
    assert got.shape == want.shape == (25,)
    assert np.allclose(got, want, rtol=1e-12, atol=1e-14), \
        np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300))


def test_spec_survey_probe_values():
    """Values recorded in SURVEY.md 8(c) for the live reference."""
    from phd_qmclib_b200 import model
    s = model.Spec(100, 1, 1, 100, 100, 25)
    assert s.obf_params.param_e0 == pytest.approx(19.511385442428583, rel=1e-13)
    assert s.tbf_params.param_k2 == pytest.approx(0.03072541267251028, rel=1e-12)
    assert s.tbf_params.param_beta == pytest.approx(0.7914901227054671, rel=1e-12)
    assert s.tbf_params.param_r_off == pytest.approx(47.14364443123984, rel=1e-12)
    assert s.tbf_params.param_am == pytest.approx(0.9778196524753014, rel=1e-12)


def test_spec_validation():
    from phd_qmclib_b200 import model
    with pytest.raises(ValueError):
        model.Spec(1, 1, 1, 10, 10, 6)          # r_m > L/2
    with pytest.raises(ValueError):
        model.Spec(1, 1, 1, 10, 10, 2, num_defects=3)
    with pytest.raises(ValueError):
        model.Spec(1, 1, 1, 10, 10, 2, num_defects=2, defect_magnitude=5)
    s = model.Spec(0, 1, 0, 4, 4, 1)
    assert s.is_free and s.is_ideal
    np.random.seed(0)
    c = s.init_get_sys_conf()
    assert c.shape == (2, 4) and np.all((c[0] >= 0) & (c[0] < 4))
    r = s.init_get_sys_conf(model.DIST_REGULAR)
    assert np.array_equal(r[0], np.arange(4.0))


def test_dmc_sampling_surface():
    from phd_qmclib_b200 import dmc, model
    spec = model.Spec(5 * math.pi ** 2, 1, 2, 8, 8, 2)
    smp = dmc.Sampling(spec, 1e-3, 64, 48)
    assert smp.num_walkers_control_factor == 0.125      # reference default
    assert isinstance(smp.rng_seed, int)
    assert smp.ddf_params.sigma_spread == math.sqrt(2e-3)
    assert smp.density_params.assume_none and smp.ssf_params.assume_none
    assert smp.state_confs_shape == (64, 2, 8)
    with pytest.raises(TypeError):
        smp.ssf_momenta
    with pytest.raises(TypeError):
        smp.density_bins_edges
    smp = dmc.Sampling(spec, 1e-3, 64, 48,
                       ssf_est_spec=dmc.SSFEstSpec(5),
                       density_est_spec=dmc.DensityEstSpec(16, False, None))
    assert np.allclose(smp.ssf_momenta, np.arange(5) * 2 * math.pi / 8)
    assert len(smp.density_bins_edges) == 17
    assert smp.density_params == dmc.DensityParams(16, False, 99999999, False)
    assert smp.cfc_spec.ssf_params.pfw_num_time_steps == 99999999
    pd = smp.core_funcs.init_props_data_block((3, 4))
    assert pd.num_walkers.dtype == np.uint64 and pd.energy.shape == (3, 4)
    assert dmc.State._fields == (
        'confs', 'props', 'energy', 'weight', 'num_walkers', 'ref_energy',
        'accum_energy', 'max_num_walkers', 'branching_spec')
    assert dmc.SamplingBlock._fields == ('iter_props', 'iter_density',
                                         'iter_ssf', 'last_state')


def test_vmc_sampling_surface():
    from phd_qmclib_b200 import model, vmc
    spec = model.Spec(5 * math.pi ** 2, 1, 2, 8, 8, 2)
    smp = vmc.Sampling(spec, 0.125, rng_seed=3)
    assert smp.tpf_params == vmc.TPFParams(8, 0.125, 0.0, 8.0)
    assert smp.ssf_params.assume_none
    with pytest.raises(TypeError):
        smp.ssf_momenta
    smp = vmc.Sampling(spec, 0.125, 3, vmc.SSFEstSpec(6))
    assert len(smp.ssf_momenta) == 6
    assert vmc.SamplingBlock._fields == ('iter_props', 'iter_ssf',
                                         'accept_rate', 'last_state')
    assert smp.core_funcs.init_props_data_block((2,)).move_stat.dtype == bool


def test_cswf_optimizer_host_surface():
    """Spec derivation for the trial cutoffs of the optimiser matches the
    reference's attr.evolve + Spec.tbf_params (the 25 scalars frozen in the
    cswf fixtures); no device needed."""
    from conftest import golden, golden_names
    from phd_qmclib_b200 import model
    for name in golden_names('cswf_'):
        g = golden(name)
        kw = dict(zip([str(k) for k in g['spec_keys']], g['spec_vals']))
        spec = model.Spec(**kw)
        opt = model.CSWFOptimizer(spec, g['confs'], g['ini_lnpsi'])
        lo, hi = opt.principal_function_bounds[0]
        assert lo == 5e-2 and hi == (0.5 - 5e-3) * spec.supercell_size
        for rm, block in zip(g['cutoffs'], g['trial_params']):
            trial = opt.update_spec(rm)
            assert trial.tbf_contact_cutoff == rm
            got = model.param_block(trial)
            assert np.max(np.abs(got - block) / np.maximum(np.abs(block), 1e-300)) < 1e-10
        v = model.CSWFOptimizer.weighed_variance(
            2 * (g['lnpsi'][0] - g['ini_lnpsi']), g['energy'][0])
        assert abs(v - g['variance'][0]) <= 1e-12 * abs(g['variance'][0])
