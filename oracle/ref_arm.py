"""TEST / BENCH INFRASTRUCTURE ONLY -- times the LIVE reference (Numba) on the
host cores, for ``bench.py --impl reference`` and its ``cpu_baseline`` leg.

The reference tree is looked for (see ``refshim.REFERENCE_SRC``) under
``$QMCB_REFERENCE_SRC``, ``/root/reference/src`` (builder container) and
``baseline/_ref`` (the offline install made by ``baseline/install_reference.sh``;
git-ignored, travels to the GPU box).  Protocol: SURVEY.md 8(d) "CPU baseline
beside it" / BASELINE.md section 3 -- ``mrbp_qmc.dmc.Sampling(...).blocks()``
(``mrbp_qmc/dmc.py:144-334``, ``qmc_base/dmc.py:815-971``), first block
discarded (JIT compilation), the following blocks timed with
``time.perf_counter``.
"""
import os
import time

import numpy as np


def probe():
    """(ok, reason): can the Numba reference run in this process?"""
    try:
        import numba  # noqa: F401
    except Exception as exc:        # pragma: no cover - image dependent
        return False, f'numba not importable ({exc.__class__.__name__}: {exc})'
    import refshim
    if not refshim.available():
        return False, ('no reference tree under $QMCB_REFERENCE_SRC, '
                       '/root/reference/src or baseline/_ref')
    return True, ''


def _load():
    import refshim
    return refshim.load()


def lattice_confs(nw, nop, seed):
    """The bench's initial population (bench.py:initial_confs)."""
    rng = np.random.default_rng(seed)
    ini = np.zeros((nw, 2, nop))
    ini[:, 0] = (np.arange(nop)[None, :] + 0.25
                 + 0.15 * (rng.random((nw, nop)) - 0.5))
    return ini


class DmcRun:
    """One reference DMC sampling kept alive: `step()` times one more block
    of `Sampling.blocks()` (the JIT compilation and the first block happen
    in the constructor and are not timed)."""

    def __init__(self, spec_kwargs, *, nw, cap, dt, nwc, nts, parallel,
                 seed=7, num_modes=0, num_bins=0):
        import numba
        mrbp = _load()
        model, dmc = mrbp.model, mrbp.dmc
        spec = model.Spec(**spec_kwargs)
        nop = spec.boson_number
        kw = {}
        if num_modes:
            kw['ssf_est_spec'] = dmc.SSFEstSpec(num_modes, as_pure_est=True,
                                                pfw_num_time_steps=nts)
        if num_bins:
            kw['density_est_spec'] = dmc.DensityEstSpec(
                num_bins, as_pure_est=True, pfw_num_time_steps=nts)
        sampling = dmc.Sampling(spec, dt, cap, nw,
                                num_walkers_control_factor=nwc, rng_seed=seed,
                                jit_parallel=bool(parallel), **kw)
        ini_state = sampling.build_state(lattice_confs(nw, nop, 11))
        self.blocks = sampling.blocks(ini_state, nts, 0)
        t0 = time.perf_counter()
        next(self.blocks)                   # JIT compilation + first block
        self.first_block_s = time.perf_counter() - t0
        self.parallel = bool(parallel)
        self.nsteps = 0
        self.cores = int(numba.get_num_threads()) if parallel else 1
        self.layer = None
        if parallel:
            try:
                self.layer = numba.threading_layer()
            except Exception:       # no parallel region has run
                self.layer = 'unknown'
        self.numba = numba.__version__
        self._what = (f'{nw} target / {cap} capacity walkers x {nts} time '
                      f'steps')
        self._est = (f', pure S(k) M={num_modes} and density B={num_bins} '
                     f'every step' if (num_modes or num_bins) else '')

    def step(self):
        """(units, seconds) of one more block."""
        t0 = time.perf_counter()
        blk = next(self.blocks)
        secs = time.perf_counter() - t0
        self.nsteps += 1
        ws = float(np.asarray(blk.iter_props.num_walkers, dtype=np.float64)
                   .sum())
        return ws, secs

    @property
    def sample(self):
        return (f'{self._what}{self._est} per bench step (first block '
                f'discarded: JIT), mrbp_qmc.dmc.Sampling(jit_parallel='
                f'{self.parallel}).blocks() of the unmodified reference '
                f'under oracle/refshim.py, numba {self.numba}')


class VmcRun:
    """Single Metropolis chain of the reference (`mrbp_qmc.vmc.Sampling`,
    `qmc_base/vmc.py:670-770`): one core by construction."""

    def __init__(self, spec_kwargs, *, move_spread, ns, num_modes, seed=1):
        import numba
        mrbp = _load()
        model, vmc = mrbp.model, mrbp.vmc
        spec = model.Spec(**spec_kwargs)
        kw = {}
        if num_modes:
            kw['ssf_est_spec'] = vmc.SSFEstSpec(num_modes)
        sampling = vmc.Sampling(spec, move_spread, rng_seed=seed, **kw)
        ini_state = sampling.build_state(
            lattice_confs(1, spec.boson_number, 0)[0])
        self.blocks = sampling.blocks(ns, ini_state)
        t0 = time.perf_counter()
        next(self.blocks)
        self.first_block_s = time.perf_counter() - t0
        self.ns = ns
        self.cores, self.layer, self.parallel = 1, None, False
        self.numba = numba.__version__

    def step(self):
        t0 = time.perf_counter()
        next(self.blocks)
        return float(self.ns), time.perf_counter() - t0

    @property
    def sample(self):
        return (f'one chain x {self.ns} steps per bench step (first block '
                f'discarded: JIT), mrbp_qmc.vmc.Sampling.blocks() of the '
                f'unmodified reference under oracle/refshim.py, numba '
                f'{self.numba}')


def cpu_model():
    try:
        for line in open('/proc/cpuinfo'):
            if line.startswith('model name'):
                return line.split(':', 1)[1].strip()
    except OSError:
        pass
    return None
