"""Blocking analysis (Flyvbjerg-Petersen) for the statistical parity tests:
the standard error of a correlated series is the plateau of the naive error
under repeated pairwise averaging."""
import numpy as np


def blocked_error(x, min_blocks=8):
    x = np.asarray(x, dtype=np.float64)
    best = x.std(ddof=1) / np.sqrt(len(x))
    while len(x) // 2 >= min_blocks:
        x = 0.5 * (x[0:2 * (len(x) // 2):2] + x[1:2 * (len(x) // 2):2])
        best = max(best, x.std(ddof=1) / np.sqrt(len(x)))
    return float(best)


def ratio_mean_error(num, den):
    """mean(num)/mean(den) with a blocked delta-method error, as the
    reference's EnergyBlocks.mean_error does in spirit
    (qmc_exec/data/dmc.py:30-75)."""
    num, den = np.asarray(num, float), np.asarray(den, float)
    r = num.mean() / den.mean()
    lin = (num - r * den) / den.mean()
    return float(r), blocked_error(lin)
