"""DMC sampling of the multi-rods Bose gas on the B200 engine.

Host-side mirror of the reference's DMC sampler interface
(``src/phd_qmclib/qmc_base/dmc.py`` and ``src/phd_qmclib/mrbp_qmc/dmc.py``):
same class, method, attribute and NamedTuple field names, same argument
meaning and error behaviour, so that the reference's procedure layer
(``qmc_exec.dmc.Proc.exec``, reference ``qmc_exec/dmc/proc.py:136-415``) can
drive it unchanged.  Nothing here computes: every block is one
``qmcb_dmc_run_block`` call into ``libqmcb200.so``.
"""
import math
import typing as t
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .engine import Engine

__all__ = ['Sampling', 'State', 'StateProps', 'BranchingSpec', 'PropsData',
           'SamplingBlock', 'SamplingStateDataBlock', 'DensityEstSpec', 'SSFEstSpec', 'StateError',
           'DDFParams', 'DensityParams', 'SSFParams', 'CFCSpec']

_BIG_NTS = 99999999       # the reference's "very large integer" default


class StateError(ValueError):
    """Flags errors related to the handling of a DMC state
    (reference mrbp_qmc/dmc.py:50-52)."""


class StateProps(t.NamedTuple):
    """Reference qmc_base/dmc.py:103-107."""
    energy: np.ndarray
    weight: np.ndarray
    mask: np.ndarray


class BranchingSpec(t.NamedTuple):
    """Reference qmc_base/dmc.py:97-100."""
    cloning_factor: np.ndarray
    cloning_ref: np.ndarray


class State(t.NamedTuple):
    """A DMC state (reference qmc_base/dmc.py:117-127)."""
    confs: np.ndarray
    props: StateProps
    energy: float
    weight: float
    num_walkers: int
    ref_energy: float
    accum_energy: float
    max_num_walkers: int
    branching_spec: t.Optional[BranchingSpec] = None


class PropsData(t.NamedTuple):
    """Per-step series of one block (reference qmc_base/dmc.py:130-143)."""
    energy: np.ndarray
    weight: np.ndarray
    num_walkers: np.ndarray
    ref_energy: np.ndarray
    accum_energy: np.ndarray


class DDFParams(t.NamedTuple):
    """Reference mrbp_qmc/dmc.py:55-61."""
    boson_number: int
    time_step: float
    sigma_spread: float
    lower_bound: float
    upper_bound: float


class DensityParams(t.NamedTuple):
    num_bins: int
    as_pure_est: bool
    pfw_num_time_steps: int
    assume_none: bool


class SSFParams(t.NamedTuple):
    num_modes: int
    as_pure_est: bool
    pfw_num_time_steps: int
    assume_none: bool


class CFCSpec(t.NamedTuple):
    model_params: tuple
    obf_params: tuple
    tbf_params: tuple
    ddf_params: DDFParams
    density_params: DensityParams
    ssf_params: SSFParams


@dataclass(frozen=True)
class DensityEstSpec:
    """Density estimator spec (reference mrbp_qmc/dmc.py:103-121)."""
    num_bins: int
    as_pure_est: bool = True
    pfw_num_time_steps: t.Optional[int] = _BIG_NTS

    def __post_init__(self):
        object.__setattr__(self, 'num_bins', int(self.num_bins))
        object.__setattr__(self, 'as_pure_est', bool(self.as_pure_est))
        if self.pfw_num_time_steps is None:
            object.__setattr__(self, 'pfw_num_time_steps', _BIG_NTS)


@dataclass(frozen=True)
class SSFEstSpec:
    """Structure factor estimator spec (reference mrbp_qmc/dmc.py:124-140)."""
    num_modes: int
    as_pure_est: bool = True
    pfw_num_time_steps: t.Optional[int] = None

    def __post_init__(self):
        if self.pfw_num_time_steps is None:
            object.__setattr__(self, 'pfw_num_time_steps', _BIG_NTS)


class SamplingBlock:
    """One block of DMC steps (reference qmc_base/dmc.py:146-152).

    Same fields as the reference NamedTuple.  ``last_state`` is fetched from
    the GPU on first access: the procedure layer reads it for the final block
    only (reference qmc_exec/dmc/proc.py:356-362), and a State is a
    capacity-sized copy of the population.  It must be read before the block
    iterator is advanced again (else :class:`StateError`).
    """
    __slots__ = ('iter_props', 'iter_density', 'iter_ssf', '_fetch', '_state')
    _fields = ('iter_props', 'iter_density', 'iter_ssf', 'last_state')

    def __init__(self, iter_props, iter_density, iter_ssf, fetch_state):
        self.iter_props = iter_props
        self.iter_density = iter_density
        self.iter_ssf = iter_ssf
        self._fetch = fetch_state
        self._state = None

    @property
    def last_state(self) -> State:
        if self._state is None:
            self._state = self._fetch()
        return self._state

    def __iter__(self):
        return iter((self.iter_props, self.iter_density, self.iter_ssf,
                     self.last_state))

    def __len__(self):
        return 4

    def __getitem__(self, i):
        return tuple(self)[i]


class SamplingStateDataBlock(t.NamedTuple):
    """Reference qmc_base/dmc.py:155-159."""
    confs: np.ndarray
    props: StateProps
    iter_props: PropsData


class CoreFuncs:
    """The few ``core_funcs`` members the procedure layer touches
    (reference qmc_exec/dmc/proc.py:208-209)."""

    @staticmethod
    def init_props_data_block(block_shape) -> PropsData:
        """Reference qmc_base/dmc.py:790-812."""
        z = lambda dt: np.zeros(block_shape, dtype=dt)   # noqa: E731
        return PropsData(z(np.float64), z(np.float64), z(np.uint64),
                         z(np.float64), z(np.float64))

    @staticmethod
    def init_branching_spec(max_num_walkers) -> BranchingSpec:
        """Reference qmc_base/dmc.py:600-611."""
        return BranchingSpec(np.zeros(max_num_walkers, dtype=np.int64),
                             np.zeros(max_num_walkers, dtype=np.int64))


core_funcs = CoreFuncs()


def local_capacity(max_num_walkers: int, world_size: int) -> int:
    """Slots per rank when ``max_num_walkers`` global slots are sharded.

    The reference truncates the population at the GLOBAL capacity only
    (qmc_base/dmc.py:636-651); a rank's share of the ensemble fluctuates more
    than the whole between two rebalances, so every slab gets a quarter more
    slots than its even share (and never fewer than 32 on top)."""
    share = -(-int(max_num_walkers) // int(world_size))
    if int(world_size) == 1:
        return share
    return share + max(32, share // 4)


def slab_bounds(n: int, world_size: int, rank: int) -> t.Tuple[int, int]:
    """[lo, hi) of rank's contiguous slab of ``n`` ordered walkers; sizes
    differ by at most one, the larger slabs first (the layout
    ``qmcb_rebalance_plan`` restores)."""
    base, extra = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_sum(dist, arrays) -> None:
    """In-place sum over the ranks of a list of float64 numpy arrays through
    an initialised ``torch.distributed`` process group (plumbing: NCCL
    reduces device tensors, gloo host tensors)."""
    import torch
    if not arrays:
        return
    flat = np.concatenate([np.asarray(a, dtype=np.float64).ravel()
                           for a in arrays])
    t_ = torch.from_numpy(flat)
    on_gpu = str(dist.get_backend()) == 'nccl'
    if on_gpu:
        t_ = t_.cuda()
    dist.all_reduce(t_)
    flat = t_.cpu().numpy()
    off = 0
    for a in arrays:
        a[...] = flat[off:off + a.size].reshape(a.shape)
        off += a.size


@dataclass(frozen=True)
class Sampling:
    """A DMC sampling (reference mrbp_qmc/dmc.py:143-334).

    ``model_spec`` is this package's :class:`~phd_qmclib_b200.model.Spec` or a
    reference ``mrbp_qmc.Spec``.  ``jit_parallel`` / ``jit_fastmath`` are
    accepted for signature compatibility and ignored.  Engine-only options:
    ``device`` (CUDA ordinal), ``energy_mode`` (0 = the reference's
    stale-slot energy in the branching weight; 1 = the parent's
    energy), ``eager_last_state`` (copy the State back after every block),
    ``dist`` (an initialised ``torch.distributed``-like module: one process
    per GPU, the walkers sharded over the ranks), ``reblock_max_order`` (keep
    on-the-fly reblocking tables of the per-step series on the device, read
    with :meth:`otf_reblock_data`).

    With ``dist``: ``max_num_walkers`` / ``target_num_walkers`` stay GLOBAL;
    rank r owns the r-th contiguous slab of the ensemble in
    ``local_capacity`` slots, so the global ensemble is the ordered
    concatenation of the slabs.  A State then holds the local
    slab (confs, props, num_walkers) next to the GLOBAL scalars (energy,
    weight, ref_energy, accum_energy); the per-step series of a block are
    global, the estimator tables are all-reduced once per block, and the
    populations are evened out by order-preserving neighbour shifts before
    each block.
    """
    model_spec: t.Any
    time_step: float
    max_num_walkers: int
    target_num_walkers: int
    num_walkers_control_factor: t.Optional[float] = None
    rng_seed: t.Optional[int] = None
    density_est_spec: t.Optional[DensityEstSpec] = None
    ssf_est_spec: t.Optional[SSFEstSpec] = None
    jit_parallel: bool = True
    jit_fastmath: bool = False
    device: int = 0
    energy_mode: int = 0
    eager_last_state: bool = False
    dist: t.Any = None
    reblock_max_order: t.Optional[int] = None
    _cache: dict = field(default_factory=dict, init=False, repr=False,
                         compare=False)

    def __post_init__(self):
        if self.rng_seed is None:
            if self.world_size > 1:
                raise ValueError('rng_seed must be given (and equal on every '
                                 'rank) when the walkers are sharded')
            seed = int(np.random.SeedSequence().generate_state(1)[0])
            object.__setattr__(self, 'rng_seed', seed)
        if self.num_walkers_control_factor is None:
            object.__setattr__(self, 'num_walkers_control_factor', 1.25e-1)

    # -- parameters, as the reference exposes them ---------------------------
    @property
    def ddf_params(self) -> DDFParams:
        z_min, z_max = self.model_spec.boundaries
        return DDFParams(self.model_spec.boson_number, self.time_step,
                         math.sqrt(2 * self.time_step), z_min, z_max)

    @property
    def density_params(self) -> DensityParams:
        s = self.density_est_spec
        if s is None:
            return DensityParams(1, False, 1, True)
        return DensityParams(s.num_bins, s.as_pure_est, s.pfw_num_time_steps,
                             False)

    @property
    def ssf_params(self) -> SSFParams:
        s = self.ssf_est_spec
        if s is None:
            return SSFParams(1, False, 1, True)
        return SSFParams(s.num_modes, s.as_pure_est, s.pfw_num_time_steps,
                         False)

    @property
    def cfc_spec(self) -> CFCSpec:
        m = self.model_spec
        return CFCSpec(m.params, m.obf_params, m.tbf_params, self.ddf_params,
                       self.density_params, self.ssf_params)

    @property
    def density_bins_edges(self) -> np.ndarray:
        if self.density_est_spec is None:
            raise TypeError('the density spec has no been specified')
        return np.linspace(0, self.model_spec.supercell_size,
                           self.density_est_spec.num_bins + 1)

    @property
    def ssf_momenta(self) -> np.ndarray:
        if self.ssf_est_spec is None:
            raise TypeError('the static structure factor spec has no been '
                            'specified')
        return (np.arange(self.ssf_est_spec.num_modes) * 2 * math.pi
                / self.model_spec.supercell_size)

    # -- sharding ---------------------------------------------------------------
    @property
    def world_size(self) -> int:
        return int(self.dist.get_world_size()) if self.dist is not None else 1

    @property
    def rank(self) -> int:
        return int(self.dist.get_rank()) if self.dist is not None else 0

    @property
    def local_capacity(self) -> int:
        """Slots on this rank."""
        return local_capacity(self.max_num_walkers, self.world_size)

    @property
    def state_confs_shape(self):
        return (self.local_capacity,) + tuple(self.model_spec.sys_conf_shape)

    @property
    def state_props_shape(self):
        return self.local_capacity,

    @property
    def core_funcs(self) -> CoreFuncs:
        return core_funcs

    # -- engine ---------------------------------------------------------------
    @property
    def engine(self) -> Engine:
        if 'engine' not in self._cache:
            eng = Engine(self.model_spec, self.device)
            if self.world_size > 1:
                eng.comm_init_torch(self.dist, self.rank, self.world_size)
            self._cache['engine'] = eng
        return self._cache['engine']

    def _engine_params(self, target_num_walkers=None):
        z_min, z_max = self.model_spec.boundaries
        dp, sp = self.density_params, self.ssf_params
        return Engine.dmc_params(
            self.time_step, self.max_num_walkers,
            target_num_walkers or self.target_num_walkers,
            self.num_walkers_control_factor, self.rng_seed, z_min, z_max,
            energy_mode=self.energy_mode,
            local_capacity=(self.local_capacity if self.world_size > 1
                            else 0),
            ssf=None if sp.assume_none else (
                sp.num_modes, sp.as_pure_est,
                min(int(sp.pfw_num_time_steps), _BIG_NTS)),
            density=None if dp.assume_none else (
                dp.num_bins, dp.as_pure_est,
                min(int(dp.pfw_num_time_steps), _BIG_NTS)))

    # -- states ----------------------------------------------------------------
    def build_state(self, sys_conf_set: np.ndarray,
                    ref_energy: float = None) -> State:
        """Initial state from a set of configurations: drift and local energy
        of the LAST ``target_num_walkers`` of them, unit weights
        (reference mrbp_qmc/dmc.py:268-328)."""
        sys_conf_set = np.asarray(sys_conf_set)
        conf_shape = tuple(self.model_spec.sys_conf_shape)
        if sys_conf_set.ndim == 3 and sys_conf_set.shape[1:] != conf_shape:
            raise StateError("sys_conf_set is not a valid set of "
                             "configurations of the model spec")
        sys_conf_set = sys_conf_set[-self.target_num_walkers:]
        world, rank = self.world_size, self.rank
        if world > 1:
            # every rank is handed the same global set and keeps its slab
            lo, hi = slab_bounds(len(sys_conf_set), world, rank)
            sys_conf_set = sys_conf_set[lo:hi]
        n, wmax = len(sys_conf_set), self.local_capacity
        if n > wmax:
            raise StateError('more configurations than max_num_walkers')
        ev = self.engine.model_eval(sys_conf_set, want=('energy', 'drift'))
        confs = np.zeros(self.state_confs_shape, dtype=np.float64)
        confs[:n, 0] = sys_conf_set[:, 0]
        confs[:n, 1] = ev['drift']
        energy = np.zeros(wmax)
        weight = np.zeros(wmax)
        mask = np.ones(wmax, dtype=bool)
        energy[:n] = ev['energy']
        weight[:n] = 1.0
        mask[:n] = False
        sums = np.array([(energy[:n] * weight[:n]).sum(), weight[:n].sum()])
        if world > 1:
            allreduce_sum(self.dist, [sums])
        state_energy, state_weight = float(sums[0]), float(sums[1])
        mean_energy = state_energy / state_weight
        if ref_energy is None:
            ref_energy = mean_energy
        return State(confs, StateProps(energy, weight, mask), state_energy,
                     state_weight, n, float(ref_energy), mean_energy, wmax,
                     core_funcs.init_branching_spec(wmax))

    def _load_state(self, ini_state: State, target_num_walkers=None):
        """Ship ``ini_state`` to the GPU as the population to branch from
        (the reference copies it into its three buffers and restarts the
        running totals, qmc_base/dmc.py:700-735)."""
        props = ini_state.props
        live = ~np.asarray(props.mask, dtype=bool)
        n = int(ini_state.num_walkers)
        if int(live.sum()) != n or not live[:n].all():
            raise StateError('the live walkers of a state must occupy its '
                             'first num_walkers slots')
        if len(props.energy) != self.local_capacity:
            raise StateError('state was built for another max_num_walkers')
        sc = _lib.StateScalars()
        sc.energy = float(ini_state.energy)
        sc.weight = float(ini_state.weight)
        sc.ref_energy = float(ini_state.ref_energy)
        sc.accum_energy = float(ini_state.accum_energy)
        sc.total_energy = 0.0
        sc.total_weight = 0.0
        sc.num_walkers = n
        sc.max_num_walkers = self.local_capacity
        sc.step = 0
        sc.capacity_hits = 0
        self.engine.dmc_set_state(
            self._engine_params(target_num_walkers),
            np.asarray(ini_state.confs)[:n],
            np.asarray(props.energy)[:n], np.asarray(props.weight)[:n], sc,
            slot_energy=np.asarray(props.energy, dtype=np.float64),
            global_slot_offset=self.rank * self.local_capacity)
        if self.reblock_max_order is not None:
            self.engine.dmc_reblock_reset(self.reblock_max_order)

    def _check_capacity(self):
        """The reference drops walkers silently when the population hits
        ``max_num_walkers`` (qmc_base/dmc.py:636-651); here the
        event is counted on the device and reported once per occurrence --
        on a sharded run it means a SLAB hit its local capacity, which the
        reference has no analogue of."""
        hits = int(self.engine.dmc_scalars().capacity_hits)
        seen = self._cache.get('capacity_hits', 0)
        if hits > seen:
            import warnings
            where = (f'rank {self.rank}: local capacity '
                     f'{self.local_capacity}' if self.world_size > 1 else
                     f'max_num_walkers = {self.max_num_walkers}')
            warnings.warn(f'DMC population truncated at capacity in '
                          f'{hits - seen} time step(s) ({where}); raise '
                          f'max_num_walkers or lower the time step',
                          RuntimeWarning, stacklevel=3)
        self._cache['capacity_hits'] = hits

    @property
    def capacity_hits(self) -> int:
        """Time steps (since the iterator started) in which branching was
        truncated at the capacity of this rank."""
        return int(self.engine.dmc_scalars().capacity_hits)

    def otf_reblock_data(self) -> t.Dict[str, np.ndarray]:
        """Accumulated on-the-fly reblocking tables of the per-step series
        (energy, weight, num_walkers, ref_energy, accum_energy) of every
        block run since the iterator started, in the reference's
        ``otf_data_dtype``: ``stats.reblock.OTFObject(table)`` gives means,
        errors per block size and the optimal block size without the series
        having left the GPU (reference ``stats/reblock.py:525-604, 652-757``).
        Needs ``reblock_max_order``."""
        if self.reblock_max_order is None:
            raise TypeError('reblock_max_order has not been specified')
        return self.engine.dmc_reblock_get()

    def _fetch_state(self, stamp=None) -> State:
        eng = self.engine
        if stamp is not None and self._cache.get('stamp') != stamp:
            raise StateError('last_state of an earlier block was not read '
                             'before the iterator advanced; read it first or '
                             'use eager_last_state=True')
        s = eng.dmc_get_state()
        sc = s['scalars']
        bs = BranchingSpec(np.zeros(self.local_capacity, dtype=np.int64),
                           s['cloning_ref'])
        return State(s['confs'],
                     StateProps(s['energy'], s['weight'],
                                s['mask'].astype(bool)),
                     float(sc.energy), float(sc.weight), int(sc.num_walkers),
                     float(sc.ref_energy), float(sc.accum_energy),
                     int(sc.max_num_walkers), bs)

    def states(self, ini_state: State) -> t.Iterator[State]:
        """Generator of DMC states, one per time step
        (reference qmc_base/dmc.py:311-323, 664-787)."""
        self._load_state(ini_state)
        while True:
            self.engine.dmc_run_block(1)
            yield self._fetch_state()

    def state_data_blocks(self, ini_state: State, num_time_steps_block: int
                          ) -> t.Iterator[SamplingStateDataBlock]:
        """Blocks that keep every State of every step: confs
        (nts, Wmax, 2, N) and props (nts, Wmax).  Like the reference, the
        population-control target of this iterator is the initial state's
        walker count (qmc_base/dmc.py:1009), not ``target_num_walkers``."""
        nts = int(num_time_steps_block)
        self._load_state(ini_state, int(ini_state.num_walkers))
        wmax = self.local_capacity
        while True:
            confs = np.zeros((nts,) + self.state_confs_shape)
            energy = np.zeros((nts, wmax))
            weight = np.zeros((nts, wmax))
            mask = np.ones((nts, wmax), dtype=bool)
            props = core_funcs.init_props_data_block((nts,))
            for i in range(nts):
                self.engine.dmc_run_block(1)
                st = self._fetch_state()
                confs[i], energy[i] = st.confs, st.props.energy
                weight[i], mask[i] = st.props.weight, st.props.mask
                props.energy[i], props.weight[i] = st.energy, st.weight
                props.num_walkers[i] = st.num_walkers
                props.ref_energy[i] = st.ref_energy
                props.accum_energy[i] = st.accum_energy
            yield SamplingStateDataBlock(
                confs, StateProps(energy, weight, mask), props)

    def blocks(self, ini_state: State, num_time_steps_blocks: int,
               burn_in_blocks: int) -> t.Iterator[SamplingBlock]:
        """Infinite generator of blocks of ``num_time_steps_blocks`` steps;
        the estimators are skipped for the first ``burn_in_blocks`` blocks
        (reference qmc_base/dmc.py:325-345, 815-971)."""
        nts = int(num_time_steps_blocks)
        dp, sp = self.density_params, self.ssf_params
        self._load_state(ini_state)
        eng = self.engine
        block_idx = 0
        while True:
            props = core_funcs.init_props_data_block((nts,))
            out = dict(zip(PropsData._fields, props))
            est = block_idx >= burn_in_blocks
            # disabled estimators: the reference's dummy arrays
            # (mrbp_qmc/dmc.py:572-579, 619-626)
            i_den = np.zeros((1, 1, 1) if dp.assume_none
                             else (nts, dp.num_bins, 1))
            i_ssf = np.zeros((1, 1, 3) if sp.assume_none
                             else (nts, sp.num_modes, 3))
            if self.world_size > 1 and block_idx > 0:
                # even out the slabs; done here, not after the block, so
                # that last_state of the previous block stays readable
                eng.dmc_rebalance()
            eng.dmc_run_block(
                nts, eval_estimators=est, out=out,
                density=None if dp.assume_none else i_den,
                ssf=None if sp.assume_none else i_ssf)
            # (sharded: the engine has summed the estimator tables over the
            # ranks on the device, over its own NCCL communicator)
            self._check_capacity()
            stamp = self._cache['stamp'] = object()
            blk = SamplingBlock(props, i_den, i_ssf,
                                lambda s=stamp: self._fetch_state(s))
            if self.eager_last_state:
                blk.last_state      # noqa: B018  (materialise now)
            yield blk
            block_idx += 1
