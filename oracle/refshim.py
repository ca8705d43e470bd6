"""TEST INFRASTRUCTURE ONLY -- loader for the live reference.

Makes the *unmodified* reference tree (``/root/reference/src``) importable on
Python 3.12 / numpy 2.x / numba 0.65 so that ``oracle/make_golden.py`` can run
the reference's own Numba code and freeze its outputs under ``tests/golden``.
Nothing of the reference is copied: this module only patches the interpreter
environment (missing third-party modules, removed aliases) and pre-seeds four
``cached_property`` slots whose bodies numba >= 0.5x can no longer type
(``np.array(tuple(namedtuple))``, reference ``mrbp_qmc/model.py:582`` and
``mrbp_qmc/dmc.py:363,378,393``).

``/root/reference`` only exists in the builder container; on the GPU box the
tree is the offline install ``baseline/_ref`` (``baseline/install_reference.sh``),
used by ``bench.py --impl reference`` / the ``cpu_baseline`` leg through
``oracle/ref_arm.py`` and by the drop-in GPU test, which skips without it.
"""
import collections
import collections.abc
import functools
import logging
import os
import sys
import types
import typing

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_reference():
    """$QMCB_REFERENCE_SRC, the builder container's tree, or the offline
    install under baseline/_ref (baseline/install_reference.sh: unmodified
    files, git-ignored, travels to the GPU box)."""
    cands = [os.environ.get('QMCB_REFERENCE_SRC'), '/root/reference/src',
             os.path.join(_ROOT, 'baseline', '_ref')]
    if os.environ.get('QMCB_REFERENCE_ONLY_ENV'):      # tests of the fallback
        cands = cands[:1]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, 'phd_qmclib')):
            return c
    return cands[0] or cands[1]


REFERENCE_SRC = _find_reference()

_installed = False


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, 'phd_qmclib'))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, n):
        return _Dummy()

    def __call__(self, *a, **k):
        return _Dummy()


def install():
    """Patch the environment; idempotent."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_SRC}')

    # py3.7 silently ignored extra bases of typing.NamedTuple classes.
    orig_new = typing.NamedTupleMeta.__new__

    def nt_new(cls, typename, bases, ns):
        bases = tuple(b for b in bases
                      if b is typing._NamedTuple or b is typing.Generic)
        return orig_new(cls, typename, bases, ns)

    typing.NamedTupleMeta.__new__ = nt_new

    # Aliases removed from numpy / collections.
    if not hasattr(np, 'int'):
        np.int = int
    if not hasattr(np, 'alltrue'):
        np.alltrue = np.all
    for n in ('Mapping', 'Sequence', 'MutableMapping'):
        if not hasattr(collections, n):
            setattr(collections, n, getattr(collections.abc, n))

    # numba must be imported before the colorama stub exists.
    import numba
    import numba.core.config as nbcfg
    import numba.core.runtime as nbrt
    sys.modules['numba.config'] = nbcfg
    numba.config = nbcfg
    sys.modules['numba.runtime'] = nbrt

    # Third-party modules the reference imports eagerly but the hot path
    # never uses.
    _mod('cached_property', cached_property=functools.cached_property)

    class _Cfg:
        def set(self, **kw):
            import contextlib
            return contextlib.nullcontext()

    dask = _mod('dask', config=_Cfg())
    dask.bag = _mod('dask.bag', from_sequence=lambda s: s)
    try:
        import h5py  # noqa: F401
    except ImportError:
        _mod('h5py', File=_Dummy, Group=_Dummy)
    _mod('colorlog', ColoredFormatter=lambda fmt, **k: logging.Formatter(
        '%(asctime)s %(name)s %(levelname)s: %(message)s'))
    ru = _mod('ruamel')
    ru.yaml = _mod('ruamel.yaml', YAML=_Dummy)
    _mod('colored', attr=lambda *a: '', fg=lambda *a: '',
         stylize=lambda s, *a: s)
    _mod('tzlocal', get_localzone=lambda: None)
    _mod('colorama', deinit=lambda: None, init=lambda *a, **k: None)
    for name in ('matplotlib', 'matplotlib.pyplot'):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                _mod(name)

    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    _installed = True


def patch_numba065():
    """Pre-seed the params->array helpers numba 0.65 cannot compile."""
    import numba as nb
    from phd_qmclib.mrbp_qmc import dmc, model

    @nb.njit
    def model_params_transform(p):
        out = np.empty(12, np.float64)
        out[0] = p.lattice_depth
        out[1] = p.lattice_ratio
        out[2] = p.interaction_strength
        out[3] = p.boson_number
        out[4] = p.supercell_size
        out[5] = p.tbf_contact_cutoff
        out[6] = p.defect_magnitude
        out[7] = p.defects_sep
        out[8] = p.well_width
        out[9] = p.barrier_width
        out[10] = p.is_free
        out[11] = p.is_ideal
        return out

    model.core_funcs.__dict__['model_params_transform'] = \
        model_params_transform

    @nb.njit
    def ddf_transform(p):
        out = np.empty(5, np.float64)
        out[0] = p.boson_number
        out[1] = p.time_step
        out[2] = p.sigma_spread
        out[3] = p.lower_bound
        out[4] = p.upper_bound
        return out

    @nb.njit
    def est_transform(p):
        out = np.empty(4, np.float64)
        out[0] = p[0]
        out[1] = p[1]
        out[2] = p[2]
        out[3] = p[3]
        return out

    for cf in list(dmc.core_funcs_table.values()):
        cf.__dict__['ddf_params_transform'] = ddf_transform
        cf.__dict__['density_params_transform'] = est_transform
        cf.__dict__['ssf_params_transform'] = est_transform


def load():
    """Return the reference's ``mrbp_qmc`` package, ready to run."""
    install()
    from phd_qmclib import mrbp_qmc
    patch_numba065()
    return mrbp_qmc
