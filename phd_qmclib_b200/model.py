"""Host-side model spec of the multi-rods Bijl-Jastrow (mrbp_qmc) Bose gas.

Mirrors the interface of the reference's ``mrbp_qmc.model.Spec``
(reference ``src/phd_qmclib/mrbp_qmc/model.py:135-400``): same attribute and
property names, same NamedTuple field names, same units (hbar^2/2m = 1,
lattice period = 1).  Only the parameter derivation lives here -- everything
that is evaluated per configuration runs on the GPU (``csrc/``).

The engine accepts either this ``Spec`` or a reference ``Spec`` instance
(duck-typed through ``params`` / ``obf_params`` / ``tbf_params``); see
:func:`param_block`.
"""
import enum
import math
import typing as t
from dataclasses import dataclass, field

import numpy as np

__all__ = ['Spec', 'Params', 'OBFParams', 'TBFParams', 'CFCSpec',
           'SysConfSlot', 'SysConfDistType', 'DIST_RAND', 'DIST_REGULAR',
           'kp_ground_state_energy', 'param_block', 'NUM_PARAMS']


class SysConfSlot(enum.IntEnum):
    """Rows of a system configuration (reference jastrow/model.py:30-38)."""
    pos = 0
    drift = 1


class SysConfDistType(enum.Enum):
    RANDOM = 'random'
    REGULAR = 'regular'


DIST_RAND = SysConfDistType.RANDOM
DIST_REGULAR = SysConfDistType.REGULAR


class Params(t.NamedTuple):
    """Reference ``mrbp_qmc.model.Params`` (model.py:40-54)."""
    lattice_depth: float
    lattice_ratio: float
    interaction_strength: float
    boson_number: int
    supercell_size: float
    tbf_contact_cutoff: float
    defect_magnitude: float
    defects_sep: int
    well_width: float
    barrier_width: float
    is_free: bool
    is_ideal: bool


class OBFParams(t.NamedTuple):
    """Reference ``OBFParams`` (model.py:57-65)."""
    lattice_depth: float
    lattice_ratio: float
    well_width: float
    barrier_width: float
    param_e0: float
    param_k1: float
    param_kp1: float


class TBFParams(t.NamedTuple):
    """Reference ``TBFParams`` (model.py:68-75)."""
    supercell_size: float
    tbf_contact_cutoff: float
    param_k2: float
    param_beta: float
    param_r_off: float
    param_am: float


class CFCSpec(t.NamedTuple):
    model_params: Params
    obf_params: OBFParams
    tbf_params: TBFParams


def _kp_dispersion(v0, r, ez, ctx):
    """Kronig-Penney band condition at zero quasi-momentum, f(E) = 0.

    Well width a = 1/(1+r), barrier width b = r/(1+r), barrier height v0
    (reference ``ideal.py:8-55``).
    """
    a = 1 / (1 + r)
    b = r / (1 + r)
    if ez == 0:
        s = ctx.sqrt(v0)
        return s * ctx.sinh(b * s) / (2 * (1 + r)) + ctx.cosh(b * s) - 1
    if ez == v0:
        s = ctx.sqrt(v0)
        return -r * s / (2 * (1 + r)) * ctx.sin(a * s) + ctx.cos(a * s) - 1
    k = ctx.sqrt(ez)
    kp = ctx.sqrt(v0 - ez)
    return ((v0 - 2 * ez) / (2 * k * kp) * ctx.sinh(b * kp) * ctx.sin(a * k)
            + ctx.cosh(b * kp) * ctx.cos(a * k) - 1)


def kp_ground_state_energy(lattice_depth: float, lattice_ratio: float) -> float:
    """Single-particle ground-state energy e0 of the KP lattice.

    Same two-stage solve as the reference (``ideal.py:58-85``): a bracketed
    double-precision root, then a polish with ``mpmath``.
    """
    import mpmath as mp
    from scipy.optimize import brentq
    v0, r = float(lattice_depth), float(lattice_ratio)
    if v0 == 0.0:
        return 0.0
    hi = min(v0, (1 + r) ** 2 * math.pi ** 2)
    try:
        root = brentq(lambda e: _kp_dispersion(v0, r, e, math), 0, hi)
        root = mp.findroot(lambda e: _kp_dispersion(v0, r, e, mp), root,
                           verify=False)
    except OverflowError:
        root = mp.findroot(lambda e: _kp_dispersion(v0, r, e, mp),
                           (0, min(v0, (1 + r) ** 2 * mp.pi ** 2)),
                           solver='illinois', verify=False)
    return float(mp.chop(root))


@dataclass(frozen=True)
class Spec:
    """Parameters of the multi-rods Bose gas with a Bijl-Jastrow trial
    function.  Field order and meaning as in the reference ``Spec``."""
    lattice_depth: float
    lattice_ratio: float
    interaction_strength: float
    boson_number: int
    supercell_size: float
    tbf_contact_cutoff: float
    num_defects: t.Optional[int] = None
    defect_magnitude: t.Optional[float] = None
    _cache: dict = field(default_factory=dict, init=False, repr=False,
                         compare=False)

    def __post_init__(self):
        set_ = lambda k, v: object.__setattr__(self, k, v)  # noqa: E731
        set_('lattice_depth', float(self.lattice_depth))
        set_('lattice_ratio', float(self.lattice_ratio))
        set_('interaction_strength', float(self.interaction_strength))
        if int(self.boson_number) != self.boson_number:
            raise ValueError('boson_number must be an integer')
        set_('boson_number', int(self.boson_number))
        set_('supercell_size', float(self.supercell_size))
        set_('tbf_contact_cutoff', float(self.tbf_contact_cutoff))
        if not abs(self.tbf_contact_cutoff) <= abs(self.supercell_size / 2):
            raise ValueError("parameter value 'rm' out of domain")
        nd, dm = self.num_defects, self.defect_magnitude
        if nd is not None:
            nd = int(nd)
            if nd < 0:
                raise ValueError("number of defects can't be negative")
            sites = int(math.ceil(self.supercell_size))
            if nd and sites % nd:
                raise ValueError(
                    f"the specified number of defects ({nd:d}) can't be "
                    f"evenly distributed in the lattice")
        if dm is not None:
            dm = float(dm)
        # Same resolution rules as the reference __attrs_post_init__
        # (model.py:172-196).
        if dm is None and nd is None:
            dm, nd = self.lattice_depth, 0
        else:
            if nd is None:
                nd, dm = 0, self.lattice_depth
            else:
                dm = dm if nd else self.lattice_depth
            if dm > self.lattice_depth:
                raise ValueError("Defect magnitude can't be greater than "
                                 "the lattice depth.")
        set_('num_defects', nd)
        set_('defect_magnitude', dm)

    # -- geometry ---------------------------------------------------------
    @property
    def boundaries(self):
        return 0., 1. * self.supercell_size

    @property
    def well_width(self):
        return 1 / (1 + self.lattice_ratio)

    @property
    def barrier_width(self):
        return self.lattice_ratio / (1 + self.lattice_ratio)

    @property
    def is_free(self):
        return self.lattice_depth <= 1e-10 or self.lattice_ratio <= 1e-10

    @property
    def is_ideal(self):
        return self.interaction_strength <= 1e-10

    @property
    def sys_conf_shape(self):
        return len(SysConfSlot), self.boson_number

    def get_sys_conf_buffer(self):
        return np.zeros(self.sys_conf_shape, dtype=np.float64)

    def init_get_sys_conf(self, dist_type=DIST_RAND, offset=None):
        """A configuration with positions laid out as ``dist_type`` says
        (reference model.py:248-273; uses numpy's global RNG like it)."""
        nop, size = self.boson_number, self.supercell_size
        z_min, _ = self.boundaries
        conf = self.get_sys_conf_buffer()
        offset = offset or 0.
        if dist_type is DIST_RAND or getattr(dist_type, 'value', None) == 'random':
            spread = size * np.random.random_sample(nop)
        elif (dist_type is DIST_REGULAR
              or getattr(dist_type, 'value', None) == 'regular'):
            spread = np.linspace(0, size, nop, endpoint=False)
        else:
            raise ValueError("unrecognized '{}' dist_type".format(dist_type))
        conf[SysConfSlot.pos, :] = z_min + (offset + spread) % size
        return conf

    # -- derived parameters ----------------------------------------------
    @property
    def params(self) -> Params:
        sites = int(math.ceil(self.supercell_size))
        nd = self.num_defects
        defects_sep = 1 if not nd else int(sites // nd)
        return Params(self.lattice_depth, self.lattice_ratio,
                      self.interaction_strength, self.boson_number,
                      self.supercell_size, self.tbf_contact_cutoff,
                      self.defect_magnitude, defects_sep, self.well_width,
                      self.barrier_width, self.is_free, self.is_ideal)

    @property
    def obf_params(self) -> OBFParams:
        if 'obf' not in self._cache:
            v0 = self.lattice_depth
            e0 = kp_ground_state_energy(v0, self.lattice_ratio)
            self._cache['obf'] = OBFParams(
                v0, self.lattice_ratio, self.well_width, self.barrier_width,
                param_e0=e0, param_k1=math.sqrt(e0),
                param_kp1=math.sqrt(v0 - e0))
        return self._cache['obf']

    @property
    def tbf_params(self) -> TBFParams:
        """Matching of the short-range cos branch to the phonon tail
        sin(pi r / L)^beta at r_m (reference model.py:318-393)."""
        if 'tbf' in self._cache:
            return self._cache['tbf']
        from scipy.optimize import brentq
        gn, nop = self.interaction_strength, self.boson_number
        size, rm = self.supercell_size, self.tbf_contact_cutoff
        if not abs(rm) <= abs(size / 2):
            raise ValueError("parameter value 'rm' out of domain")
        if gn == 0:
            out = TBFParams(size, rm, param_k2=0., param_beta=0.,
                            param_r_off=1 / 2 * size, param_am=1.0)
            self._cache['tbf'] = out
            return out
        pi, tan, sin, cos = math.pi, math.tan, math.sin, math.cos
        lieb_gamma = 0.5 * (size / nop) ** 2 * gn
        a1d = 2.0 / (lieb_gamma * nop)      # 1D scattering length, box units
        x = rm / size                       # r_m in box units

        def beta_x(u):
            # beta * x as a function of u = k2 * r_m
            if u == 0:
                return tan(pi * x) / pi
            return (u / pi * (x - u * a1d * tan(u)) * tan(pi * x)
                    / (u * a1d + x * tan(u)))

        def mismatch(u):
            # continuity of the local energy at r_m
            bx = beta_x(u)
            return ((u * sin(pi * x)) ** 2 + (pi * bx * cos(pi * x)) ** 2
                    - pi ** 2 * bx * x)

        u = brentq(mismatch, 0, pi / 2)
        bx = (u / pi * (x - u * a1d * tan(u)) * tan(pi * x)
              / (u * a1d + x * tan(u)))
        k2 = u / x
        k2_r_off = math.atan(1 / (k2 * a1d))
        beta = bx / x
        r_off = k2_r_off / k2
        am = sin(pi * x) ** beta / cos(u - k2_r_off)
        out = TBFParams(size, rm, param_k2=k2 / size, param_beta=beta,
                        param_r_off=r_off * size, param_am=am)
        self._cache['tbf'] = out
        return out

    @property
    def cfc_spec(self) -> CFCSpec:
        return CFCSpec(self.params, self.obf_params, self.tbf_params)


NUM_PARAMS = 25


def param_block(spec) -> np.ndarray:
    """Flatten a model spec into the 25-double block of ``qmcb_model_params``
    (``include/qmcb200.h``): model(12) + obf(7) + tbf(6), in the order of the
    reference's own params->array transforms (model.py:571-686).

    ``spec`` is this module's :class:`Spec`, a reference ``mrbp_qmc.Spec``, or
    a ``(params, obf_params, tbf_params)`` triple.
    """
    if hasattr(spec, 'params') and hasattr(spec, 'obf_params'):
        triple = (spec.params, spec.obf_params, spec.tbf_params)
    else:
        triple = tuple(spec)[:3]
    flat = [float(v) for part in triple for v in tuple(part)]
    if len(flat) != NUM_PARAMS:
        raise ValueError(f'expected {NUM_PARAMS} parameters, got {len(flat)}')
    return np.array(flat, dtype=np.float64)


def _evolve(spec, **changes):
    """``attr.evolve`` for a reference (attrs) Spec, ``dataclasses.replace``
    for this module's."""
    if hasattr(type(spec), '__attrs_attrs__'):
        import attr
        return attr.evolve(spec, **changes)
    from dataclasses import replace
    return replace(spec, **changes)


class PhysicalFuncs:
    """Batched physical properties of a model spec on the GPU: the
    reference's ``PhysicalFuncs`` (``mrbp_qmc/model.py:801-814``,
    gufuncs in ``qmc_base/jastrow/model.py:1007-1122``) with the same
    call signatures and NumPy broadcasting over leading dimensions.

    ``spec`` may be this module's :class:`Spec`, a reference ``Spec``, or a
    ``(params, obf_params, tbf_params)`` triple (the reference's
    ``cfc_spec_nt``).
    """

    def __init__(self, cfc_spec_nt, device: int = 0):
        self.cfc_spec_nt = cfc_spec_nt
        self._device = device
        self._engine = None

    @classmethod
    def from_model_spec(cls, model_spec, device: int = 0):
        """Reference: ``PhysicalFuncs.from_model_spec``
        (``mrbp_qmc/model.py:806-809``)."""
        return cls(model_spec.cfc_spec, device)

    @property
    def engine(self):
        if self._engine is None:
            from .engine import Engine
            self._engine = Engine(self.cfc_spec_nt, self._device)
        return self._engine

    def _batch(self, sys_conf):
        confs = np.asarray(sys_conf, dtype=np.float64)
        nop = self.engine.nop
        if confs.ndim < 2 or confs.shape[-2:] != (len(SysConfSlot), nop):
            raise ValueError(
                f'sys_conf must have shape (..., {len(SysConfSlot)}, {nop})')
        lead = confs.shape[:-2]
        return np.ascontiguousarray(confs.reshape((-1,) + confs.shape[-2:])), lead

    def wf_abs_log(self, sys_conf):
        """``(ns,nop)->()``: ln|Psi| (``jastrow/model.py:1019-1043``)."""
        confs, lead = self._batch(sys_conf)
        out = self.engine.model_eval(confs, want=('lnpsi',))['lnpsi']
        return out.reshape(lead)[()]

    def energy(self, sys_conf):
        """``(ns,nop)->()``: local energy (``jastrow/model.py:1045-1067``)."""
        confs, lead = self._batch(sys_conf)
        out = self.engine.model_eval(confs, want=('energy',))['energy']
        return out.reshape(lead)[()]

    def one_body_density(self, sz, sys_conf):
        """``(),(ns,nop)->()``: one-body density matrix estimator
        (``jastrow/model.py:1069-1091``); ``sz`` broadcasts against the
        leading dimensions of ``sys_conf``."""
        confs, lead = self._batch(sys_conf)
        sz = np.asarray(sz, dtype=np.float64)
        offsets, inv = np.unique(sz.ravel(), return_inverse=True)
        table = self.engine.one_body_density(confs, offsets)    # (B, S)
        b_idx = np.arange(confs.shape[0]).reshape(lead)
        s_idx = inv.reshape(sz.shape)
        b_idx, s_idx = np.broadcast_arrays(b_idx, s_idx)
        return table[b_idx, s_idx][()]

    def fourier_density(self, kz_set, sys_conf):
        """``(nkz),(ns,nop)->(nkz)``: complex rho_k
        (``jastrow/model.py:1093-1122``)."""
        confs, lead = self._batch(sys_conf)
        kz_set = np.asarray(kz_set, dtype=np.float64)
        if kz_set.ndim != 1:
            raise ValueError('kz_set must be one-dimensional')
        out = self.engine.fourier_density_k(confs, kz_set)
        return out.reshape(lead + (kz_set.shape[0],))


class CSWFOptimizer:
    """Correlated-sampling optimiser of the trial wave function: minimises
    the weighted variance of the local energy over a fixed configuration set
    as a function of ``tbf_contact_cutoff``.  Interface of the reference's
    ``CSWFOptimizer`` (``mrbp_qmc/model.py:817-942``,
    ``qmc_base/jastrow/model.py:1125-1211``); the configuration set stays on
    the GPU and each trial value costs one model-evaluation launch plus one
    reduction launch instead of a dask bag over the configurations.
    ``use_threads`` / ``num_workers`` are accepted for compatibility and
    ignored.
    """

    def __init__(self, spec, sys_conf_set, ini_wf_abs_log_set,
                 ref_energy=None, use_threads=True, num_workers=None,
                 verbose=False, device: int = 0):
        self.spec = spec
        self.sys_conf_set = np.asarray(sys_conf_set, dtype=np.float64)
        self.ini_wf_abs_log_set = np.asarray(ini_wf_abs_log_set,
                                             dtype=np.float64)
        self.ref_energy = ref_energy
        self.use_threads = use_threads
        self.num_workers = num_workers
        self.verbose = verbose
        self._device = device
        self._engine = None

    @property
    def engine(self):
        if self._engine is None:
            from .engine import Engine
            self._engine = Engine(self.spec, self._device)
            self._engine.cs_load(self.sys_conf_set, self.ini_wf_abs_log_set)
        return self._engine

    def update_spec(self, tbf_contact_cutoff: float):
        """Reference: ``mrbp_qmc/model.py:852-862``."""
        return _evolve(self.spec,
                       tbf_contact_cutoff=float(tbf_contact_cutoff))

    @staticmethod
    def weighed_variance(weights_log_set, energy_set, ref_energy=None):
        """Variance of E_L under the weights exp(weights_log_set), about their
        weighted mean (API of ``jastrow/model.py:1147-1165``; like there,
        ``ref_energy`` is accepted and not used).  Host-side helper for
        arrays already on the host -- the optimiser's objective is reduced on
        the device by ``qmcb_cs_variance`` (``cs_variance_kernel``)."""
        lw = np.asarray(weights_log_set, dtype=np.float64).ravel()
        e = np.asarray(energy_set, dtype=np.float64).ravel()
        w = np.exp(lw - lw.max())       # largest weight is 1: no overflow
        w /= w.sum()
        mean = np.dot(w, e)
        return float(np.dot(w, np.square(e - mean)))

    def wf_abs_log_and_energy_set(self, cfc_spec):
        """ln|Psi| and E_L of every configuration under ``cfc_spec``
        (``mrbp_qmc/model.py:889-903``)."""
        res = self.engine.cs_variance(cfc_spec, want_sets=True)
        return res['wf_abs_log'], res['energy']

    def principal_function(self, tbf_contact_cutoff):
        """The weighted variance at a trial cutoff
        (``jastrow/model.py:1186-1206``)."""
        tbf_contact_cutoff = float(np.ravel(tbf_contact_cutoff)[0])
        trial = self.update_spec(tbf_contact_cutoff)
        return self.engine.cs_variance(trial)['variance']

    @property
    def principal_function_bounds(self):
        """Reference: ``mrbp_qmc/model.py:905-914``."""
        sc_size = self.spec.supercell_size
        return [(5e-2, (0.5 - 5e-3) * sc_size)]

    def exec(self, **de_kwargs):
        """Differential evolution over the cutoff on the host
        (``mrbp_qmc/model.py:929-942``); returns the updated spec."""
        from scipy.optimize import differential_evolution
        opt = differential_evolution(self.principal_function,
                                     bounds=self.principal_function_bounds,
                                     disp=self.verbose, **de_kwargs)
        opt_cutoff, = opt.x
        return self.update_spec(opt_cutoff)
