"""GPU drop-in test: the UNMODIFIED reference procedure layer
(`qmc_exec.dmc.Proc.exec`, reference qmc_exec/dmc/proc.py:136-415, and
`qmc_exec.vmc.Proc.exec`, qmc_exec/vmc/proc.py:87-250) drives the REAL engine
through the binding of INTEGRATION.md.  The reference tree is the offline
install under baseline/_ref (baseline/install_reference.sh) or
/root/reference in the builder container, imported through oracle/refshim.py;
skipped, with the reason, where neither exists."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _b200_proc import make_dmc_proc, make_vmc_proc  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ref():
    import refshim
    if not refshim.available():
        pytest.skip('no reference tree (baseline/_ref, /root/reference/src or '
                    '$QMCB_REFERENCE_SRC)')
    try:
        import numba  # noqa: F401  (the reference imports it at module level)
    except ImportError as exc:
        pytest.skip(f'numba not importable: {exc}')
    return refshim.load()


def test_reference_dmc_proc_exec_on_the_gpu(ref):
    from phd_qmclib.mrbp_qmc import Spec, dmc_exec
    from phd_qmclib_b200 import dmc as b200_dmc
    Proc = make_dmc_proc(dmc_exec, b200_dmc)
    nop = 16
    spec = Spec(lattice_depth=5 * np.pi ** 2, lattice_ratio=1,
                interaction_strength=2, boson_number=nop, supercell_size=nop,
                tbf_contact_cutoff=0.25 * nop)
    nblk, nts = 6, 32
    proc = Proc(spec, 1e-3, 640, 512, rng_seed=5, num_blocks=nblk,
                num_time_steps_block=nts, burn_in_blocks=2,
                density_spec=dmc_exec.DensityEstSpec(num_bins=64),
                ssf_spec=dmc_exec.SSFEstSpec(num_modes=8))
    np.random.seed(3)
    proc_input = dmc_exec.ProcInput.from_model_sys_conf_spec(
        dmc_exec.ModelSysConfSpec(dist_type='REGULAR'), proc)
    assert type(proc_input.state).__name__ == 'State'
    result = proc.exec(proc_input)
    from phd_qmclib_b200._lib import LIB_PATH
    assert os.path.exists(LIB_PATH)             # the CUDA library did the work
    blocks = result.data.blocks
    assert blocks.energy.totals.shape == (nblk,)
    e = blocks.energy.mean / nop
    assert 14.0 < e < 16.5                      # E/N of this lattice gas ~15.1
    assert blocks.density.totals.shape == (nblk, 64)
    assert blocks.ss_factor.fdk_sqr_abs_part.totals.shape == (nblk, 8)
    # pure estimators report the last step of a block (reduce_fac,
    # qmc_exec/dmc/proc.py:319-320): k = 0 gives N^2 per live walker
    sk = blocks.ss_factor.fdk_sqr_abs_part
    assert np.allclose(sk.totals[:, 0],
                       nop ** 2 * np.asarray(sk.weight_totals)[:, 0],
                       rtol=1e-12)
    # pure density: a running average of N counts per live walker
    den_per_walker = blocks.density.totals.sum(axis=1) \
        / np.asarray(blocks.density.weight_totals)[:, 0]
    assert np.all(np.abs(den_per_walker / nop - 1) < 0.1)
    # the final State round-trips into a restart through the reference's own
    # ProcInput.from_result (mrbp_qmc/dmc_exec/proc.py:131-143)
    st = result.state
    assert st.confs.shape == (640, 2, nop) and st.props.mask.dtype == bool
    assert len(st.branching_spec) == 2
    assert int((~st.props.mask).sum()) == st.num_walkers
    again = proc.exec(dmc_exec.ProcInput.from_result(result, proc))
    assert again.data.blocks.energy.totals.shape == (nblk,)
    assert abs(again.data.blocks.energy.mean / nop - e) < 0.3
    proc.sampling.engine.close()


def test_reference_vmc_proc_exec_on_the_gpu(ref):
    from phd_qmclib.mrbp_qmc import Spec, vmc_exec
    from phd_qmclib_b200 import vmc as b200_vmc
    Proc = make_vmc_proc(vmc_exec, b200_vmc)
    spec = Spec(lattice_depth=5 * np.pi ** 2, lattice_ratio=1,
                interaction_strength=4, boson_number=16, supercell_size=16,
                tbf_contact_cutoff=4)
    proc = Proc(spec, 0.25 * spec.well_width, rng_seed=1, num_blocks=6,
                num_steps_block=256, burn_in_blocks=2,
                ssf_spec=vmc_exec.SSFEstSpec(num_modes=8))
    np.random.seed(0)
    proc_input = vmc_exec.ProcInput.from_model_sys_conf_spec(
        vmc_exec.ModelSysConfSpec(dist_type='REGULAR'), proc)
    result = proc.exec(proc_input)
    blocks = result.data.blocks
    assert blocks.energy.totals.shape == (6,)
    assert 14.0 < blocks.energy.mean / 16 < 19.0
    assert blocks.ss_factor.fdk_sqr_abs_part.totals.shape == (6, 8)
    # block means of |rho_0|^2 = N^2
    assert np.allclose(blocks.ss_factor.fdk_sqr_abs_part.totals[:, 0],
                       16.0 ** 2, rtol=1e-12)
    assert result.state.sys_conf.shape == (2, 16)
