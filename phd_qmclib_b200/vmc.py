"""VMC sampling of the multi-rods Bose gas on the B200 engine.

Host-side mirror of the reference's VMC sampler interface
(``src/phd_qmclib/qmc_base/vmc.py`` and ``src/phd_qmclib/mrbp_qmc/vmc.py``):
same class, method, attribute and NamedTuple field names.  A reference
``Sampling`` advances ONE Metropolis chain; this one advances a batch of
independent chains, each of which is the reference's single-chain algorithm
(``qmc_base/vmc.py:557-648``).  ``build_state`` of one ``(2, N)``
configuration gives the reference's single-chain shapes; a ``(C, 2, N)``
batch gives arrays with a leading chain axis.
"""
import math
import typing as t
from dataclasses import dataclass, field

import numpy as np

from .engine import Engine

__all__ = ['Sampling', 'State', 'PropsData', 'SamplingBlock',
           'SamplingStateDataBlock', 'SSFEstSpec', 'StateError', 'TPFParams',
           'SSFParams', 'CFCSpec', 'STAT_REJECTED', 'STAT_ACCEPTED']

STAT_REJECTED = 0
STAT_ACCEPTED = 1


class StateError(ValueError):
    """Flags errors related to the handling of a VMC state
    (reference mrbp_qmc/vmc.py:56-58)."""


class State(t.NamedTuple):
    """Reference qmc_base/vmc.py:128-132."""
    sys_conf: np.ndarray
    wf_abs_log: t.Union[float, np.ndarray]
    move_stat: t.Union[int, np.ndarray]


class PropsData(t.NamedTuple):
    """Reference qmc_base/vmc.py:135-139."""
    wf_abs_log: np.ndarray
    energy: np.ndarray
    move_stat: np.ndarray


class SamplingBlock(t.NamedTuple):
    """Reference qmc_base/vmc.py:142-147."""
    iter_props: PropsData
    iter_ssf: np.ndarray
    accept_rate: t.Union[float, np.ndarray]
    last_state: t.Optional[State] = None


class SamplingStateDataBlock(t.NamedTuple):
    """Reference qmc_base/vmc.py:150-155."""
    confs: np.ndarray
    props: PropsData
    accept_rate: t.Union[float, np.ndarray]
    last_state: t.Optional[State] = None


class TPFParams(t.NamedTuple):
    """Reference mrbp_qmc/vmc.py:29-34."""
    boson_number: int
    move_spread: float
    lower_bound: float
    upper_bound: float


class SSFParams(t.NamedTuple):
    num_modes: int
    supercell_size: float
    assume_none: bool


class CFCSpec(t.NamedTuple):
    model_params: tuple
    obf_params: tuple
    tbf_params: tuple
    tpf_params: TPFParams
    ssf_params: SSFParams


@dataclass
class SSFEstSpec:
    """Structure factor estimator spec (reference mrbp_qmc/vmc.py:61-66)."""
    num_modes: int


class CoreFuncs:
    """``core_funcs`` members the procedure layer touches
    (reference qmc_exec/vmc/proc.py:153-154)."""

    @staticmethod
    def init_props_data_block(block_shape) -> PropsData:
        return PropsData(np.zeros(block_shape, dtype=np.float64),
                         np.zeros(block_shape, dtype=np.float64),
                         np.zeros(block_shape, dtype=np.bool_))


core_funcs = CoreFuncs()


def chains_before(dist, num_chains: int) -> int:
    """Number of chains on the ranks below this one (plumbing through an
    initialised ``torch.distributed`` process group)."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    t_ = torch.zeros(world, dtype=torch.int64)
    t_[rank] = int(num_chains)
    if str(dist.get_backend()) == 'nccl':
        t_ = t_.cuda()
    dist.all_reduce(t_)
    return int(t_[:rank].sum().item())


@dataclass(frozen=True)
class Sampling:
    """The spec of a VMC sampling (reference mrbp_qmc/vmc.py:69-170).

    Engine-only options: ``device``; ``chain_offset`` (global index of the
    first chain of this sampling: chains are keyed by it in the counter-based
    RNG); ``dist`` (an initialised ``torch.distributed``-like module: every
    rank runs its own batch of chains -- they never interact, so there is no
    data-path collective -- and ``chain_offset`` becomes the number of chains
    on the lower ranks, which needs ``rng_seed`` to be given and equal on all
    ranks)."""
    model_spec: t.Any
    move_spread: float
    rng_seed: t.Optional[int] = None
    ssf_est_spec: t.Optional[SSFEstSpec] = None
    device: int = 0
    chain_offset: int = 0
    dist: t.Any = None
    _cache: dict = field(default_factory=dict, init=False, repr=False,
                         compare=False)

    def __post_init__(self):
        if self.rng_seed is None:
            if self.dist is not None and self.dist.get_world_size() > 1:
                raise ValueError('rng_seed must be given (and equal on every '
                                 'rank) when the chains are spread over '
                                 'ranks')
            seed = int(np.random.SeedSequence().generate_state(1)[0])
            object.__setattr__(self, 'rng_seed', seed)

    def _chain_offset(self, num_chains: int) -> int:
        """Global index of this rank's first chain: the chains of the lower
        ranks come first."""
        if self.dist is None or self.dist.get_world_size() == 1:
            return int(self.chain_offset)
        return int(self.chain_offset) + chains_before(self.dist, num_chains)

    @property
    def tpf_params(self) -> TPFParams:
        z_min, z_max = self.model_spec.boundaries
        return TPFParams(self.model_spec.boson_number, self.move_spread,
                         z_min, z_max)

    @property
    def ssf_params(self) -> SSFParams:
        size = self.model_spec.supercell_size
        if self.ssf_est_spec is None:
            return SSFParams(1, size, True)
        return SSFParams(self.ssf_est_spec.num_modes, size, False)

    @property
    def cfc_spec(self) -> CFCSpec:
        m = self.model_spec
        return CFCSpec(m.params, m.obf_params, m.tbf_params, self.tpf_params,
                       self.ssf_params)

    @property
    def ssf_momenta(self) -> np.ndarray:
        if self.ssf_est_spec is None:
            raise TypeError('the static structure factor spec has no been '
                            'specified')
        return (np.arange(self.ssf_est_spec.num_modes) * 2 * math.pi
                / self.model_spec.supercell_size)

    @property
    def core_funcs(self) -> CoreFuncs:
        return core_funcs

    @property
    def engine(self) -> Engine:
        if 'engine' not in self._cache:
            self._cache['engine'] = Engine(self.model_spec, self.device)
        return self._cache['engine']

    # -- states ----------------------------------------------------------------
    def build_state(self, sys_conf: np.ndarray) -> State:
        """``sys_conf`` (2, N) -> the reference's single-chain State;
        (C, 2, N) -> a batch of C chains (reference mrbp_qmc/vmc.py:145-170).
        """
        sys_conf = np.asarray(sys_conf, dtype=np.float64)
        shape = tuple(self.model_spec.sys_conf_shape)
        if sys_conf.shape != shape and sys_conf.shape[1:] != shape:
            raise StateError("sys_conf is not a valid configuration "
                             "of the model spec")
        ln = self.engine.model_eval(sys_conf, want=('lnpsi',))['lnpsi']
        if sys_conf.ndim == 2:
            return State(sys_conf, float(ln[0]), STAT_ACCEPTED)
        return State(sys_conf, ln,
                     np.full(len(ln), STAT_ACCEPTED, dtype=np.int64))

    #: 0 = uniform proposal of width ``move_spread``; the gaussian variant
    #: (``vmc_ndf.Sampling``) overrides both
    _proposal = 0

    def _proposal_scale(self):
        return self.move_spread

    def _start(self, ini_state: State):
        conf = np.asarray(ini_state.sys_conf, dtype=np.float64)
        single = conf.ndim == 2
        z_min, z_max = self.model_spec.boundaries
        sp = self.ssf_params
        self.engine.vmc_init(conf[None] if single else conf,
                             self._proposal_scale(), self.rng_seed, z_min,
                             z_max,
                             ssf_num_modes=0 if sp.assume_none
                             else sp.num_modes,
                             chain_offset=self._chain_offset(
                                 1 if single else len(conf)),
                             proposal=self._proposal)
        return single

    def _last_state(self, single, stat_last) -> State:
        confs, ln = self.engine.vmc_get_state()
        if single:
            return State(confs[0], float(ln[0]), int(stat_last[0]))
        return State(confs, ln, stat_last.astype(np.int64))

    def blocks(self, num_steps_block: int,
               ini_state: State) -> t.Iterator[SamplingBlock]:
        """Infinite generator of blocks of ``num_steps_block`` states; the
        first state of the first block is the initial one, flagged ACCEPTED
        (reference qmc_base/vmc.py:231-242, 616-618, 670-770)."""
        ns = int(num_steps_block)
        single = self._start(ini_state)
        sp = self.ssf_params
        while True:
            o = self.engine.vmc_run_block(ns, series=True)
            stat = o['move_stat'].astype(np.bool_)
            ssf = o['ssf'] if not sp.assume_none else None
            if single:
                props = PropsData(o['lnpsi'][0], o['energy'][0], stat[0])
                i_ssf = ssf[0] if ssf is not None else np.zeros((1, 1, 3))
                acc = float(o['accept_rate'][0])
            else:
                props = PropsData(o['lnpsi'], o['energy'], stat)
                i_ssf = ssf if ssf is not None else np.zeros((1, 1, 3))
                acc = o['accept_rate']
            yield SamplingBlock(props, i_ssf, acc,
                                self._last_state(single, stat[:, -1]))

    def block_sums(self, num_steps_block: int, ini_state: State):
        """Batched-run helper with no per-step series: yields dicts of the
        per-chain block sums (energy, energy^2, S(k) parts) and acceptance
        rates accumulated on the device."""
        ns = int(num_steps_block)
        self._start(ini_state)
        while True:
            yield self.engine.vmc_run_block(ns, series=False, sums=True)

    def one_body_density_blocks(self, num_steps_block: int, ini_state: State,
                                pos_offset) -> t.Iterator[np.ndarray]:
        """The one-body density matrix estimator g1(s) along the chains
        (the reference's VMC hook ``one_body_density(step_idx, pos_offset,
        sys_conf, cfc_spec, iter_obd_array)``, qmc_base/jastrow/vmc.py:267-301,
        which its own ``blocks()`` never reaches for this model): after every
        block of ``num_steps_block`` steps, g1 of the state each chain is in,
        at the displacements ``pos_offset`` -> array (chains, len(pos_offset))
        (one row for a single chain).  Evaluated on the device where the
        chains live."""
        ns = int(num_steps_block)
        single = self._start(ini_state)
        pos_offset = np.atleast_1d(np.asarray(pos_offset, dtype=np.float64))
        while True:
            self.engine.vmc_run_block(ns, series=False)
            obd = self.engine.vmc_one_body_density(pos_offset)
            yield obd[0] if single else obd

    def states(self, ini_state: State) -> t.Iterator[State]:
        """One State per Metropolis step (reference qmc_base/vmc.py:244-253).
        """
        single = self._start(ini_state)
        while True:
            o = self.engine.vmc_run_block(1, series=True)
            yield self._last_state(single, o['move_stat'][:, -1])

    def state_data_blocks(self, num_steps_block: int, ini_state: State
                          ) -> t.Iterator[SamplingStateDataBlock]:
        """Blocks that keep every configuration
        (reference qmc_base/vmc.py:848-902)."""
        ns = int(num_steps_block)
        single = self._start(ini_state)
        nop = self.model_spec.boson_number
        nch = 1 if single else len(ini_state.sys_conf)
        while True:
            yield self._chain_block(ns, single, nch, nop)

    def _chain_block(self, ns, single, nch, nop):
        # one launch: the kernel records every state of the chains on the
        # device (qmcb_vmc_run_chain)
        o = self.engine.vmc_run_chain(ns)
        confs, ln, en = o['confs'], o['lnpsi'], o['energy']
        stat = o['move_stat'].astype(np.bool_)
        acc = stat.mean(axis=1)
        last = self._last_state(single, stat[:, -1])
        if single:
            return SamplingStateDataBlock(
                confs[0], PropsData(ln[0], en[0], stat[0]), float(acc[0]),
                last)
        return SamplingStateDataBlock(confs, PropsData(ln, en, stat), acc,
                                      last)

    def as_chain(self, num_steps: int,
                 ini_state: State) -> SamplingStateDataBlock:
        """The chain with every configuration kept, e.g. to seed a DMC run
        (reference qmc_base/vmc.py:215-229, 773-902)."""
        ns = int(num_steps)
        single = self._start(ini_state)
        nop = self.model_spec.boson_number
        nch = 1 if single else len(ini_state.sys_conf)
        return self._chain_block(ns, single, nch, nop)
