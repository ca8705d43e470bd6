"""CPU: the broadcasting of the PhysicalFuncs mirror against the gufunc
signatures of the reference (qmc_base/jastrow/model.py:1007-1122), with the
engine replaced by an oracle-backed test double (this container has no GPU).

The reference's own gufuncs do not compile: they call the core functions as
``f(sys_conf, cfc_spec_nt)`` while those take the three parameter tuples
separately (numba: "missing a required argument: 'obf_params'", v0.17.0 with
numba 0.65 here), so the expected values are formed by broadcasting the live
reference's scalar core functions by hand over the same signatures."""
import numpy as np
import pytest

from specs import SPECS


class OracleEngine:
    """The four Engine calls PhysicalFuncs makes, answered by the oracle."""

    def __init__(self, spec, device=0):
        import oracle
        from phd_qmclib_b200.model import param_block
        self.o, self.p = oracle, param_block(spec)
        self.nop = int(self.p[3])

    def model_eval(self, confs, want=('lnpsi', 'energy', 'drift')):
        return self.o.model_eval(self.p, np.asarray(confs), want=want)

    def one_body_density(self, confs, offsets):
        return self.o.one_body_density(self.p, confs, offsets)

    def fourier_density_k(self, confs, kz_set):
        return self.o.fourier_density_k(self.p, confs, kz_set)

    def cs_load(self, sys_conf_set, ini_wf_abs_log_set=None):
        self.cs = (np.asarray(sys_conf_set), np.asarray(ini_wf_abs_log_set))

    def cs_variance(self, trial_spec=None, want_sets=False):
        from phd_qmclib_b200.model import param_block
        p = self.p if trial_spec is None else param_block(trial_spec)
        o = self.o.model_eval(p, self.cs[0], want=('lnpsi', 'energy'))
        var, eref = self.o.weighed_variance(o['lnpsi'], self.cs[1],
                                            o['energy'])
        out = dict(variance=var, ref_energy=eref)
        if want_sets:
            out.update(wf_abs_log=o['lnpsi'], energy=o['energy'])
        return out


@pytest.fixture(scope='module')
def ref():
    import refshim
    return refshim.load()


@pytest.mark.parametrize('name', ['ll_n16', 'defects_n20', 'odd_n7'])
def test_physical_funcs_broadcast_like_the_reference(ref, name, monkeypatch):
    from phd_qmclib_b200 import engine as engine_mod, model
    monkeypatch.setattr(engine_mod, 'Engine', OracleEngine)
    rspec = ref.model.Spec(**SPECS[name])
    cf, cfc = ref.model.core_funcs, rspec.cfc_spec
    with pytest.raises(Exception):           # the reference's gufunc is broken
        ref.model.PhysicalFuncs.from_model_spec(rspec).energy(
            rspec.init_get_sys_conf())
    # the mirror takes the reference's Spec as is
    pf = model.PhysicalFuncs.from_model_spec(rspec)
    nop, size = rspec.boson_number, rspec.supercell_size
    rng = np.random.default_rng(5)
    confs = np.zeros((3, 4, 2, nop))
    confs[..., 0, :] = rng.random((3, 4, nop)) * size

    def over_confs(fn):
        return np.array([[fn(c) for c in row] for row in confs])

    # (ns,nop)->(): leading dimensions are kept, a single configuration
    # gives a scalar
    for name_ in ('wf_abs_log', 'energy'):
        want = over_confs(lambda c: getattr(cf, name_)(c, *cfc))
        got = getattr(pf, name_)(confs)
        assert got.shape == want.shape == (3, 4)
        assert np.allclose(got, want, rtol=1e-12, atol=0)
        assert np.ndim(getattr(pf, name_)(confs[0, 0])) == 0
    # (),(ns,nop)->(): offsets broadcast against the leading dimensions
    sz = np.array([0.0, 0.4, -1.3, 0.5 * size])
    want = np.array([over_confs(lambda c: cf.one_body_density(s_, c, *cfc))
                     for s_ in sz])
    got = pf.one_body_density(sz[:, None, None], confs)
    assert got.shape == want.shape == (4, 3, 4)
    assert np.allclose(got, want, rtol=1e-12, atol=0)
    got = pf.one_body_density(sz, confs[0])      # (4,) with (4, 2, N)
    want = np.array([cf.one_body_density(s_, c, *cfc)
                     for s_, c in zip(sz, confs[0])])
    assert got.shape == (4,) and np.allclose(got, want, rtol=1e-12, atol=0)
    # (nkz),(ns,nop)->(nkz)
    kz = np.array([0.0, 2 * np.pi / size, 0.77, -3.1])
    want = np.array([[[cf.fourier_density(k, c, *cfc) for k in kz]
                      for c in row] for row in confs])
    got = pf.fourier_density(kz, confs)
    assert got.shape == want.shape == (3, 4, 4)
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)


def test_cs_optimizer_mirror_on_the_oracle(monkeypatch):
    """Host logic of CSWFOptimizer (update_spec -> trial parameters ->
    objective) with the oracle standing in for the engine, against the
    objective values of the live reference frozen in tests/golden/."""
    from conftest import golden, golden_names
    from phd_qmclib_b200 import engine as engine_mod, model
    monkeypatch.setattr(engine_mod, 'Engine', OracleEngine)
    for name in golden_names('cswf_'):
        g = golden(name)
        kw = dict(zip([str(k) for k in g['spec_keys']], g['spec_vals']))
        opt = model.CSWFOptimizer(model.Spec(**kw), g['confs'],
                                  g['ini_lnpsi'])
        for k, rm in enumerate(g['cutoffs']):
            scale = max(g['variance'][k],
                        1e-12 * np.mean(g['energy'][k] ** 2))
            assert abs(opt.principal_function(rm) - g['variance'][k]) \
                < 1e-9 * scale
            # scipy hands the objective a length-1 array
            assert opt.principal_function(np.array([rm])) \
                == opt.principal_function(rm)
