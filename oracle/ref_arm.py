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


def dmc_walker_steps_per_s(spec_kwargs, *, nw, cap, dt, nwc, nts, parallel,
                           budget_s=20.0, max_blocks=6, seed=7,
                           num_modes=0, num_bins=0):
    """Run the reference's DMC ``blocks()``; returns a dict with the
    throughput (walker-steps/s over the timed blocks), the thread count, the
    Numba threading layer and a description of the sample."""
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
        kw['density_est_spec'] = dmc.DensityEstSpec(num_bins,
                                                    as_pure_est=True,
                                                    pfw_num_time_steps=nts)
    sampling = dmc.Sampling(spec, dt, cap, nw,
                            num_walkers_control_factor=nwc, rng_seed=seed,
                            jit_parallel=bool(parallel), **kw)
    ini_state = sampling.build_state(lattice_confs(nw, nop, 11))
    blocks = sampling.blocks(ini_state, nts, 0)
    t0 = time.perf_counter()
    next(blocks)                        # JIT compilation + first block
    jit_s = time.perf_counter() - t0
    ws, secs, nb = 0.0, 0.0, 0
    while nb < max_blocks and (nb < 1 or secs < budget_s):
        t0 = time.perf_counter()
        blk = next(blocks)
        secs += time.perf_counter() - t0
        ws += float(np.asarray(blk.iter_props.num_walkers, dtype=np.float64)
                    .sum())
        nb += 1
    threads = numba.get_num_threads() if parallel else 1
    layer = None
    if parallel:
        try:
            layer = numba.threading_layer()
        except Exception:       # no parallel region has run
            layer = 'unknown'
    return dict(value=ws / secs, cores=int(threads), threading_layer=layer,
                blocks=nb, seconds=secs, first_block_s=jit_s,
                numba=numba.__version__,
                sample=(f'{nw} target / {cap} capacity walkers x {nts} time '
                        f'steps x {nb} blocks (first block discarded: JIT), '
                        f'mrbp_qmc.dmc.Sampling(jit_parallel={bool(parallel)})'
                        f'.blocks() of the unmodified reference under '
                        f'oracle/refshim.py, numba {numba.__version__}'))


def vmc_chain_steps_per_s(spec_kwargs, *, move_spread, ns, num_modes,
                          budget_s=15.0, max_blocks=8, seed=1):
    """Single Metropolis chain of the reference (``mrbp_qmc.vmc.Sampling``,
    ``qmc_base/vmc.py:670-770``): chain-steps/s on one core."""
    import numba
    mrbp = _load()
    model, vmc = mrbp.model, mrbp.vmc
    spec = model.Spec(**spec_kwargs)
    nop = spec.boson_number
    kw = {}
    if num_modes:
        kw['ssf_est_spec'] = vmc.SSFEstSpec(num_modes)
    sampling = vmc.Sampling(spec, move_spread, rng_seed=seed, **kw)
    ini = lattice_confs(1, nop, 0)[0]
    ini_state = sampling.build_state(ini)
    blocks = sampling.blocks(ns, ini_state)
    t0 = time.perf_counter()
    next(blocks)
    jit_s = time.perf_counter() - t0
    secs, nb = 0.0, 0
    while nb < max_blocks and (nb < 1 or secs < budget_s):
        t0 = time.perf_counter()
        next(blocks)
        secs += time.perf_counter() - t0
        nb += 1
    return dict(value=ns * nb / secs, cores=1, threading_layer=None,
                blocks=nb, seconds=secs, first_block_s=jit_s,
                numba=numba.__version__,
                sample=(f'one chain x {ns} steps x {nb} blocks (first block '
                        f'discarded: JIT), mrbp_qmc.vmc.Sampling.blocks() of '
                        f'the unmodified reference under oracle/refshim.py, '
                        f'numba {numba.__version__}'))


def cpu_model():
    try:
        for line in open('/proc/cpuinfo'):
            if line.startswith('model name'):
                return line.split(':', 1)[1].strip()
    except OSError:
        pass
    return None
