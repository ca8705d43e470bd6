"""CPU: the UNCHANGED reference procedure layer (qmc_exec.dmc.Proc.exec /
qmc_exec.vmc.Proc.exec, imported from /root/reference or baseline/_ref
through oracle/refshim.py) drives this package's sampler
classes through the injection point INTEGRATION.md documents -- a subclass
overriding the `sampling` cached_property.

There is no GPU here, so the engine behind the samplers is replaced by a TEST
DOUBLE backed by the C oracle: what is verified is the boundary (names,
shapes, dtypes, call order, State round trip), not the kernels."""
import functools
import os

import numpy as np
import pytest

import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(
    os.path.abspath(__file__))), 'oracle'))
import refshim  # noqa: E402

pytestmark = pytest.mark.skipif(not refshim.available(),
                                reason='reference tree not present')


class OracleEngine:
    """Test double with the Engine methods the samplers call."""

    def __init__(self, spec, device=0):
        import oracle
        from phd_qmclib_b200 import model
        self.o = oracle
        self.p = model.param_block(spec)
        self.nop = int(self.p[3])
        self.size = float(self.p[4])

    from phd_qmclib_b200.engine import Engine as _E
    dmc_params = staticmethod(_E.dmc_params)

    def model_eval(self, confs, want=('lnpsi', 'energy', 'drift')):
        return self.o.model_eval(self.p, confs, want=want)

    def dmc_set_state(self, params, confs, energy, weight, sc,
                      slot_energy=None, global_slot_offset=0):
        wmax, n = int(params.max_num_walkers), len(confs)
        st = self.o.DMCState(self.p, confs, wmax)
        for b in st.bufs:
            b['confs'][:n] = confs
            b['energy'][:n] = energy
            b['weight'][:n] = weight
        st.act['energy'][:] = slot_energy
        st.scal[:] = (sc.ref_energy, sc.total_energy, sc.total_weight)
        st.cnt[:] = (n, sc.step)
        self.st, self.dp = st, params
        self.last = None

    def dmc_run_block(self, nts, eval_estimators=False, out=None,
                      density=None, ssf=None):
        dp, st = self.dp, self.st
        s = d = None
        if dp.ssf_num_modes:
            s = dict(num=dp.ssf_num_modes, pure=bool(dp.ssf_as_pure),
                     pfw=dp.ssf_pfw_nts,
                     iter=np.zeros((nts, dp.ssf_num_modes, 3)),
                     aux=np.zeros((2, st.wmax, dp.ssf_num_modes, 3)))
        if dp.density_num_bins:
            d = dict(num=dp.density_num_bins, pure=bool(dp.density_as_pure),
                     pfw=dp.density_pfw_nts,
                     iter=np.zeros((nts, dp.density_num_bins)),
                     aux=np.zeros((2, st.wmax, dp.density_num_bins)))
        it = st.run_block(dp.rng_seed, dp.time_step, dp.target_num_walkers,
                          dp.nwc_factor, nts, dp.lower_bound, dp.upper_bound,
                          energy_mode=dp.energy_mode,
                          eval_est=bool(eval_estimators), ssf=s, density=d)
        for k in out:
            out[k][...] = it[k]
        if eval_estimators and ssf is not None:
            ssf[...] = s['iter']
        if eval_estimators and density is not None:
            density[...] = d['iter'][:, :, None]
        self.last = it
        return out

    def dmc_scalars(self):
        from phd_qmclib_b200 import _lib
        return _lib.StateScalars()

    def dmc_get_state(self, want_confs=True):
        from phd_qmclib_b200 import _lib
        st, it = self.st, self.last
        sc = _lib.StateScalars()
        sc.energy, sc.weight = it['energy'][-1], it['weight'][-1]
        sc.ref_energy, sc.accum_energy = (it['ref_energy'][-1],
                                          it['accum_energy'][-1])
        sc.num_walkers, sc.max_num_walkers = st.num_walkers, st.wmax
        return dict(confs=st.act['confs'].copy(),
                    energy=st.act['energy'].copy(),
                    weight=st.act['weight'].copy(),
                    mask=st.act['mask'].copy(), cloning_ref=st.ref.copy(),
                    scalars=sc)

    # -- VMC ------------------------------------------------------------
    def vmc_init(self, confs, move_spread, rng_seed, lower, upper,
                 ssf_num_modes=0, chain_offset=0, proposal=0):
        self.v = dict(cur=np.array(confs, dtype=np.float64), spread=move_spread,
                      seed=rng_seed, lo=lower, hi=upper, M=ssf_num_modes,
                      off=chain_offset, step0=0, first=True, prop=proposal)
        c = len(confs)
        self.v['ln'] = self.o.model_eval(self.p, self.v['cur'],
                                         want=('lnpsi',))['lnpsi']
        self.v['e'] = np.zeros(c)
        self.v['s'] = np.zeros((c, max(ssf_num_modes, 1), 3))

    def vmc_run_block(self, ns, series=True, sums=False):
        v = self.v
        a = self.o.vmc_block(self.p, v['seed'], v['spread'], v['lo'],
                             v['hi'], v['cur'], v['ln'], v['e'],
                             v['s'] if v['M'] else None, v['M'], ns,
                             v['step0'], v['first'], chain_offset=v['off'],
                             proposal=v['prop'])
        v['step0'] += ns - (1 if v['first'] else 0)
        v['first'] = False
        return dict(lnpsi=a['lnpsi'], energy=a['energy'],
                    move_stat=a['stat'], ssf=a['ssf'],
                    accept_rate=a['accept_rate'])

    def vmc_get_state(self):
        return self.v['cur'].copy(), self.v['ln'].copy()


@pytest.fixture(scope='module')
def ref():
    import refshim
    mrbp = refshim.load()
    return mrbp


@pytest.fixture()
def doubled(monkeypatch):
    from phd_qmclib_b200 import dmc, vmc
    monkeypatch.setattr(dmc, 'Engine', OracleEngine)
    monkeypatch.setattr(vmc, 'Engine', OracleEngine)
    return dmc, vmc


def test_reference_dmc_proc_drives_b200_sampling(ref, doubled):
    import attr
    from phd_qmclib.mrbp_qmc import Spec, dmc_exec
    b200_dmc, _ = doubled

    @attr.s(auto_attribs=True, frozen=True)
    class B200Proc(dmc_exec.Proc):
        """The binding of INTEGRATION.md: only `sampling` is overridden."""

        @functools.cached_property
        def sampling(self):
            nts = self.num_time_steps_block
            den = ssf = None
            if self.should_eval_density:
                den = b200_dmc.DensityEstSpec(self.density_spec.num_bins,
                                              self.density_spec.as_pure_est,
                                              nts)
            if self.should_eval_ssf:
                ssf = b200_dmc.SSFEstSpec(self.ssf_spec.num_modes,
                                          self.ssf_spec.as_pure_est, nts)
            return b200_dmc.Sampling(
                self.model_spec, self.time_step, self.max_num_walkers,
                self.target_num_walkers, self.num_walkers_control_factor,
                self.rng_seed, density_est_spec=den, ssf_est_spec=ssf)

    spec = Spec(lattice_depth=5 * np.pi ** 2, lattice_ratio=1,
                interaction_strength=2, boson_number=8, supercell_size=8,
                tbf_contact_cutoff=2)
    proc = B200Proc(spec, 1e-3, 48, 32, rng_seed=5, num_blocks=4,
                    num_time_steps_block=16, burn_in_blocks=1,
                    density_spec=dmc_exec.DensityEstSpec(num_bins=32),
                    ssf_spec=dmc_exec.SSFEstSpec(num_modes=8))
    np.random.seed(3)
    proc_input = dmc_exec.ProcInput.from_model_sys_conf_spec(
        dmc_exec.ModelSysConfSpec(dist_type='RANDOM'), proc)
    assert type(proc_input.state).__name__ == 'State'
    result = proc.exec(proc_input)
    blocks = result.data.blocks
    assert blocks.energy.totals.shape == (4,)
    e = blocks.energy.mean / 8
    assert 10 < e < 25                       # E/N of the N=8 lattice gas
    assert blocks.density.totals.shape == (4, 32)
    assert blocks.ss_factor.fdk_sqr_abs_part.totals.shape == (4, 8)
    # k = 0 mode of the pure estimator: |rho_0|^2 = N^2 per walker, so the
    # block entry is N^2 times the last step's population
    nw_last = blocks.num_walkers.totals / 16       # not exact; loose check
    assert np.all(blocks.ss_factor.fdk_sqr_abs_part.totals[:, 0] > 0)
    # the final State round-trips into a restart through the reference's own
    # ProcInput.from_result
    st = result.state
    assert st.confs.shape == (48, 2, 8) and st.props.mask.dtype == bool
    assert len(st.branching_spec) == 2
    again = proc.exec(dmc_exec.ProcInput.from_result(result, proc))
    assert again.data.blocks.energy.totals.shape == (4,)
    assert nw_last.shape == (4,)


def test_reference_hdf5_handler_round_trip(ref, doubled, monkeypatch, tmp_path):
    """the result of a B200-backed procedure goes through the
    reference's UNCHANGED HDF5FileHandler (qmc_exec/io.py:76-132,
    qmc_exec/dmc/io.py:35-98, qmc_exec/data/dmc.py hdf5_export) into the
    reference's file layout and back into a ProcResult that restarts.  h5py
    is not in this image: tests/_fake_h5py.py stands in for the dozen calls
    the handler makes."""
    import attr
    import _fake_h5py
    from phd_qmclib.mrbp_qmc import Spec, dmc_exec
    from phd_qmclib.qmc_exec import io as io_base
    from phd_qmclib.qmc_exec.dmc import io as dmc_io
    from phd_qmclib.qmc_exec.data import dmc as data_dmc
    for mod in (io_base, dmc_io, data_dmc):
        monkeypatch.setattr(mod, 'h5py', _fake_h5py, raising=False)
    b200_dmc, _ = doubled

    @attr.s(auto_attribs=True, frozen=True)
    class B200Proc(dmc_exec.Proc):
        @functools.cached_property
        def sampling(self):
            nts = self.num_time_steps_block
            return b200_dmc.Sampling(
                self.model_spec, self.time_step, self.max_num_walkers,
                self.target_num_walkers, self.num_walkers_control_factor,
                self.rng_seed,
                density_est_spec=b200_dmc.DensityEstSpec(
                    self.density_spec.num_bins,
                    self.density_spec.as_pure_est, nts),
                ssf_est_spec=b200_dmc.SSFEstSpec(
                    self.ssf_spec.num_modes, self.ssf_spec.as_pure_est, nts))

    spec = Spec(lattice_depth=5 * np.pi ** 2, lattice_ratio=1,
                interaction_strength=2, boson_number=8, supercell_size=8,
                tbf_contact_cutoff=2)
    proc = B200Proc(spec, 1e-3, 48, 32, rng_seed=5, num_blocks=3,
                    num_time_steps_block=8, burn_in_blocks=1,
                    density_spec=dmc_exec.DensityEstSpec(num_bins=16),
                    ssf_spec=dmc_exec.SSFEstSpec(num_modes=4))
    np.random.seed(3)
    result = proc.exec(dmc_exec.ProcInput.from_model_sys_conf_spec(
        dmc_exec.ModelSysConfSpec(dist_type='RANDOM'), proc))
    handler = dmc_exec.io.HDF5FileHandler(str(tmp_path / 'out.h5'), 'run-a')
    handler.dump(result)
    tree = set(_fake_h5py.File(handler.location_path, 'r').paths())
    # the layout of SURVEY 8(f) N2
    for path in ('/run-a/dmc/state/confs', '/run-a/dmc/state/branching_spec',
                 '/run-a/dmc/state/props/energy',
                 '/run-a/dmc/state/props/weight',
                 '/run-a/dmc/state/props/mask', '/run-a/dmc/proc_spec',
                 '/run-a/dmc/data/blocks/energy/totals',
                 '/run-a/dmc/data/blocks/energy/weight_totals',
                 '/run-a/dmc/data/blocks/weight/totals',
                 '/run-a/dmc/data/blocks/num_walkers/totals',
                 '/run-a/dmc/data/blocks/density/totals',
                 '/run-a/dmc/data/blocks/ss_factor/fdk_sqr_abs/totals',
                 '/run-a/dmc/data/blocks/ss_factor/fdk_real/totals',
                 '/run-a/dmc/data/blocks/ss_factor/fdk_imag/totals'):
        assert path in tree, (path, sorted(tree))
    back = handler.load()
    st, st0 = back.state, result.state
    assert np.array_equal(st.confs, st0.confs)
    assert np.array_equal(st.props.energy, st0.props.energy)
    assert np.array_equal(st.props.mask, st0.props.mask)
    assert np.array_equal(np.asarray(st.branching_spec),
                          np.asarray(st0.branching_spec))
    for k in ('energy', 'weight', 'num_walkers', 'ref_energy',
              'accum_energy', 'max_num_walkers'):
        assert getattr(st, k) == getattr(st0, k), k
    assert np.array_equal(back.data.blocks.energy.totals,
                          result.data.blocks.energy.totals)
    assert np.array_equal(back.data.blocks.density.totals,
                          result.data.blocks.density.totals)
    # and the State read back from the file restarts the B200 procedure
    again = proc.exec(dmc_exec.ProcInput.from_result(back, proc))
    assert again.data.blocks.energy.totals.shape == (3,)


def test_reference_vmc_proc_drives_b200_sampling(ref, doubled):
    import attr
    from phd_qmclib.mrbp_qmc import Spec, vmc_exec
    _, b200_vmc = doubled

    @attr.s(auto_attribs=True, frozen=True)
    class B200Proc(vmc_exec.Proc):
        @functools.cached_property
        def sampling(self):
            ssf = None
            if self.should_eval_ssf:
                ssf = b200_vmc.SSFEstSpec(self.ssf_spec.num_modes)
            return b200_vmc.Sampling(self.model_spec, self.move_spread,
                                     self.rng_seed, ssf_est_spec=ssf)

    spec = Spec(lattice_depth=5 * np.pi ** 2, lattice_ratio=1,
                interaction_strength=4, boson_number=8, supercell_size=8,
                tbf_contact_cutoff=2)
    proc = B200Proc(spec, 0.25 * spec.well_width, rng_seed=1, num_blocks=4,
                    num_steps_block=64, burn_in_blocks=1,
                    ssf_spec=vmc_exec.SSFEstSpec(num_modes=8))
    np.random.seed(0)
    proc_input = vmc_exec.ProcInput.from_model_sys_conf_spec(
        vmc_exec.ModelSysConfSpec(dist_type='RANDOM'), proc)
    result = proc.exec(proc_input)
    blocks = result.data.blocks
    assert blocks.energy.totals.shape == (4,)
    assert 10 < blocks.energy.mean / 8 < 30
    assert blocks.ss_factor.fdk_sqr_abs_part.totals.shape == (4, 8)
    assert result.state.sys_conf.shape == (2, 8)
