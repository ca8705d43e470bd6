"""Gaussian-proposal VMC sampling (reference ``qmc_base/vmc_ndf.py``,
``mrbp_qmc/vmc_ndf.py``): identical to :mod:`phd_qmclib_b200.vmc` except that
every particle is displaced by N(0, sigma^2), sigma = sqrt(time_step),
instead of a uniform step."""
import math
import typing as t
from dataclasses import dataclass, field

from . import vmc as _vmc
from .vmc import (CFCSpec, PropsData, SSFEstSpec, SSFParams,  # noqa: F401
                  SamplingBlock, SamplingStateDataBlock, State, StateError)

__all__ = ['Sampling', 'TPFParams']


class TPFParams(t.NamedTuple):
    """Reference mrbp_qmc/vmc_ndf.py:13-21."""
    boson_number: int
    sigma: float
    lower_bound: float
    upper_bound: float


@dataclass(frozen=True)
class Sampling(_vmc.Sampling):
    """The spec of the gaussian-proposal VMC sampling
    (reference mrbp_qmc/vmc_ndf.py:24-60).  ``time_step`` is the variance of
    the proposal; ``move_spread`` is inherited and unused, as in the
    reference."""
    model_spec: t.Any = None
    move_spread: float = 0.0
    rng_seed: t.Optional[int] = None
    ssf_est_spec: t.Optional[SSFEstSpec] = None
    device: int = 0
    chain_offset: int = 0
    time_step: float = None
    _cache: dict = field(default_factory=dict, init=False, repr=False,
                         compare=False)
    _proposal = 1

    def __post_init__(self):
        if self.model_spec is None or self.time_step is None:
            raise TypeError('model_spec and time_step are required')
        if not self.time_step > 0:
            raise ValueError('time_step must be positive')
        super().__post_init__()

    @property
    def tpf_params(self) -> TPFParams:
        z_min, z_max = self.model_spec.boundaries
        return TPFParams(self.model_spec.boson_number,
                         math.sqrt(self.time_step), z_min, z_max)

    def _proposal_scale(self):
        return math.sqrt(self.time_step)
