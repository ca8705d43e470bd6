"""Shared fixtures.  `gpu` marks tests that need a B200 (run by the driver on
the GPU box); everything else must pass on the CPU-only builder box."""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ORACLE_DIR = os.path.join(ROOT, 'oracle')
if ORACLE_DIR not in sys.path:
    sys.path.insert(0, ORACLE_DIR)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200)')


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def golden_names(prefix):
    return sorted(os.path.basename(p)
                  for p in glob.glob(os.path.join(GOLDEN, prefix + '*.npz')))


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def maxnorm_err(a, b):
    """Row-wise |a-b|_inf / |b|_inf (the drift criterion of SURVEY 8c).  A
    row whose own norm is tiny (the regular lattice has zero drift by
    symmetry) is measured against the typical row norm of the batch."""
    a, b = np.atleast_2d(a), np.atleast_2d(b)
    rown = np.max(np.abs(b), axis=-1)
    den = np.maximum(rown, max(float(np.median(rown)), 1e-300))
    return float(np.max(np.max(np.abs(a - b), axis=-1) / den))


def scaled_err(a, b):
    """|a-b| / max(|b|, median |b| of the batch): relative error that does
    not blow up on entries that are small by cancellation."""
    a, b = np.atleast_1d(a).astype(float), np.atleast_1d(b).astype(float)
    den = np.maximum(np.abs(b), max(float(np.median(np.abs(b))), 1e-300))
    return float(np.max(np.abs(a - b) / den))


def conditioned_rel_err(a, b, floor=1e-2):
    """Strict |a-b|/|b| over the entries that are not small by cancellation:
    |b| >= floor * median|b|.  The few entries below the floor (a sum of
    O(1e3) terms that lands within 1 % of zero) are bounded through
    ``scaled_err``; no fp64 implementation, the reference included, holds a
    relative error there."""
    a, b = np.atleast_1d(a).astype(float), np.atleast_1d(b).astype(float)
    keep = np.abs(b) >= floor * float(np.median(np.abs(b)))
    assert keep.mean() > 0.97
    return rel_err(a[keep], b[keep])


@pytest.fixture(scope='session')
def oracle():
    import oracle as _o
    _o.lib()
    return _o
