"""Thin object wrapper over the C ABI: one :class:`Engine` = one
``qmcb_handle`` = one model on one GPU.  numpy in, numpy out; no compute on
the Python side."""
import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import DMCParams, EngineError, StateScalars, VMCParams, ptr
from .model import param_block

__all__ = ['Engine', 'EngineError', 'pinned_empty', 'measure_fp64_peak']


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class _PinnedBlock:
    """Owner of one cudaHostAlloc block (freed when the last view dies)."""

    def __init__(self, nbytes):
        self._L = _lib.load()
        self.ptr = self._L.qmcb_host_alloc(nbytes)
        if not self.ptr:
            raise EngineError(f'qmcb_host_alloc({nbytes}) failed')
        self.nbytes = nbytes

    def __del__(self):
        try:
            self._L.qmcb_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float64):
    """A page-locked numpy array (for fast host<->device copies)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    block = _PinnedBlock(max(n * dtype.itemsize, 1))
    buf = (C.c_byte * block.nbytes).from_address(block.ptr)
    buf._owner = block
    a = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
    return a


def measure_fp64_peak(device=0, sustained_seconds=0.0):
    """DFMA throughput of ``device`` in TFLOP/s: burst (best single launch)
    or, with ``sustained_seconds`` > 0, back-to-back launches for that long."""
    L = _lib.load()
    tf, ms = C.c_double(), C.c_double()
    if sustained_seconds > 0:
        rc = L.qmcb_measure_fp64_sustained(device, sustained_seconds,
                                           C.byref(tf))
    else:
        rc = L.qmcb_measure_fp64_peak(device, C.byref(tf), C.byref(ms))
    if rc != 0:
        raise EngineError(f'fp64 peak measurement failed ({rc})')
    return tf.value


def comm_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    rc = _lib.load().qmcb_comm_unique_id(buf)
    if rc != 0:
        raise EngineError(f'qmcb_comm_unique_id failed ({rc}): NCCL not '
                          f'loadable')
    return bytes(buf)


class Engine:
    """The CUDA engine for one model spec on one device."""

    def __init__(self, spec, device: int = 0):
        self._L = _lib.load()
        self.block = param_block(spec)
        self.nop = int(self.block[3])
        self.supercell_size = float(self.block[4])
        self.device = int(device)
        mp = _lib.model_params_struct(self.block)
        h = C.c_void_p()
        rc = self._L.qmcb_create(C.byref(mp), self.device, C.byref(h))
        if rc != 0:
            msg = self._L.qmcb_last_error(None).decode()
            raise EngineError(f'qmcb_create failed ({rc}): {msg}')
        self._h = h
        self._dmc_cap = None

    # -- plumbing ---------------------------------------------------------
    def close(self):
        if getattr(self, '_h', None):
            self._L.qmcb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            msg = self._L.qmcb_last_error(self._h).decode()
            raise EngineError(f'{what} failed ({rc}): {msg}')

    @property
    def handle(self):
        return self._h

    @property
    def stream(self):
        """The engine's cudaStream_t as an integer."""
        return self._L.qmcb_stream(self._h) or 0

    # -- fixed-configuration evaluation ------------------------------------
    def model_eval(self, confs, want=('lnpsi', 'energy', 'drift')):
        """``confs`` (B, 2, N) or (2, N) -> dict of lnpsi (B,), energy (B,),
        drift (B, N).  Reference: ``core_funcs.wf_abs_log / energy / drift``.
        """
        confs = _f64(confs)
        if confs.ndim == 2:
            confs = confs[None]
        if confs.ndim != 3 or confs.shape[1:] != (2, self.nop):
            raise ValueError(f'confs must have shape (B, 2, {self.nop})')
        n = confs.shape[0]
        out = dict(
            lnpsi=np.empty(n) if 'lnpsi' in want else None,
            energy=np.empty(n) if 'energy' in want else None,
            drift=np.empty((n, self.nop)) if 'drift' in want else None)
        rc = self._L.qmcb_model_eval(self._h, ptr(confs), n,
                                     ptr(out['lnpsi']), ptr(out['energy']),
                                     ptr(out['drift']))
        self._check(rc, 'qmcb_model_eval')
        return out

    def model_eval_device(self, d_confs, nconf, d_lnpsi=0, d_energy=0,
                          d_drift=0):
        """Device-pointer variant (integers, e.g. ``tensor.data_ptr()``)."""
        rc = self._L.qmcb_model_eval_device(
            self._h, C.c_void_p(d_confs), nconf, C.c_void_p(d_lnpsi or None),
            C.c_void_p(d_energy or None), C.c_void_p(d_drift or None))
        self._check(rc, 'qmcb_model_eval_device')

    def fourier_density(self, confs, num_modes):
        confs = _f64(confs)
        if confs.ndim == 2:
            confs = confs[None]
        out = np.empty((confs.shape[0], num_modes, 3))
        rc = self._L.qmcb_fourier_density(self._h, ptr(confs),
                                          confs.shape[0], num_modes, ptr(out))
        self._check(rc, 'qmcb_fourier_density')
        return out

    def one_body_density(self, confs, offsets):
        """``confs`` (B, 2, N) or (2, N), ``offsets`` (S,) -> (B, S): the
        one-body density matrix estimator of each configuration at each
        displacement.  Reference: ``core_funcs.one_body_density``
        (``qmc_base/jastrow/model.py:934-965``).
        """
        confs = _f64(confs)
        if confs.ndim == 2:
            confs = confs[None]
        if confs.ndim != 3 or confs.shape[1:] != (2, self.nop):
            raise ValueError(f'confs must have shape (B, 2, {self.nop})')
        offsets = _f64(np.atleast_1d(offsets))
        if offsets.ndim != 1:
            raise ValueError('offsets must be one-dimensional')
        out = np.empty((confs.shape[0], offsets.shape[0]))
        rc = self._L.qmcb_one_body_density(
            self._h, ptr(confs), confs.shape[0], ptr(offsets),
            offsets.shape[0], ptr(out))
        self._check(rc, 'qmcb_one_body_density')
        return out

    def one_body_density_device(self, d_confs, nconf, d_offsets, num_offsets,
                                d_out):
        """Device-pointer variant, asynchronous on the engine's stream."""
        rc = self._L.qmcb_one_body_density_device(
            self._h, C.c_void_p(d_confs), nconf, C.c_void_p(d_offsets),
            num_offsets, C.c_void_p(d_out))
        self._check(rc, 'qmcb_one_body_density_device')

    def fourier_density_k(self, confs, kz_set):
        """``confs`` (B, 2, N), ``kz_set`` (K,) -> complex (B, K): rho_k at
        arbitrary momenta.  Reference: ``PhysicalFuncs.fourier_density``
        (``qmc_base/jastrow/model.py:1093-1122``)."""
        confs = _f64(confs)
        if confs.ndim == 2:
            confs = confs[None]
        if confs.ndim != 3 or confs.shape[1:] != (2, self.nop):
            raise ValueError(f'confs must have shape (B, 2, {self.nop})')
        kz = _f64(np.atleast_1d(kz_set))
        out = np.empty((confs.shape[0], kz.shape[0], 2))
        rc = self._L.qmcb_fourier_density_k(self._h, ptr(confs),
                                            confs.shape[0], ptr(kz),
                                            kz.shape[0], ptr(out))
        self._check(rc, 'qmcb_fourier_density_k')
        return out[..., 0] + 1j * out[..., 1]

    # -- trial-wave-function optimisation -----------------------------------
    def set_model_params(self, spec):
        """Swap the model scalars of the live handle (same boson number)."""
        block = param_block(spec)
        mp = _lib.model_params_struct(block)
        rc = self._L.qmcb_set_model_params(self._h, C.byref(mp))
        self._check(rc, 'qmcb_set_model_params')
        self.block = block
        self.supercell_size = float(block[4])

    def cs_load(self, sys_conf_set, ini_wf_abs_log_set=None):
        """Keep the optimiser's configuration set on the device."""
        confs = _f64(sys_conf_set)
        if confs.ndim != 3 or confs.shape[1:] != (2, self.nop):
            raise ValueError(f'confs must have shape (B, 2, {self.nop})')
        ln0 = None
        if ini_wf_abs_log_set is not None:
            ln0 = _f64(ini_wf_abs_log_set)
            if ln0.shape != (confs.shape[0],):
                raise ValueError('ini_wf_abs_log_set must have one entry per '
                                 'configuration')
        rc = self._L.qmcb_cs_load(self._h, ptr(confs), confs.shape[0],
                                  ptr(ln0))
        self._check(rc, 'qmcb_cs_load')
        self._cs_n = confs.shape[0]

    def cs_variance(self, trial_spec=None, want_sets=False):
        """Weighted variance of E_L over the loaded set under the trial
        parameters -> dict(variance, ref_energy[, wf_abs_log, energy])."""
        mp = None
        if trial_spec is not None:
            mp = C.byref(_lib.model_params_struct(param_block(trial_spec)))
        var, eref = C.c_double(), C.c_double()
        n = getattr(self, '_cs_n', 0)
        ln = np.empty(n) if want_sets else None
        en = np.empty(n) if want_sets else None
        rc = self._L.qmcb_cs_variance(self._h, mp, C.byref(var),
                                      C.byref(eref), ptr(ln), ptr(en))
        self._check(rc, 'qmcb_cs_variance')
        out = dict(variance=var.value, ref_energy=eref.value)
        if want_sets:
            out.update(wf_abs_log=ln, energy=en)
        return out

    # -- DMC ----------------------------------------------------------------
    @staticmethod
    def dmc_params(time_step, max_num_walkers, target_num_walkers,
                   nwc_factor, rng_seed, lower_bound, upper_bound,
                   energy_mode=0, ssf=None, density=None, local_capacity=0):
        """``ssf`` / ``density``: ``(num, as_pure, pfw_nts)`` or None."""
        p = DMCParams()
        p.time_step = time_step
        p.nwc_factor = nwc_factor
        p.lower_bound, p.upper_bound = lower_bound, upper_bound
        p.max_num_walkers = max_num_walkers
        p.target_num_walkers = target_num_walkers
        p.rng_seed = int(rng_seed) & 0xFFFFFFFFFFFFFFFF
        p.energy_mode = energy_mode
        if ssf:
            p.ssf_num_modes, p.ssf_as_pure, p.ssf_pfw_nts = (
                int(ssf[0]), int(bool(ssf[1])), int(ssf[2]))
        if density:
            p.density_num_bins, p.density_as_pure, p.density_pfw_nts = (
                int(density[0]), int(bool(density[1])), int(density[2]))
        p.local_capacity = local_capacity
        return p

    def dmc_init(self, params: DMCParams, ini_confs, ref_energy=None,
                 global_slot_offset=0):
        ini_confs = _f64(ini_confs)
        if ini_confs.ndim != 3 or ini_confs.shape[1:] != (2, self.nop):
            raise ValueError(f'ini_confs must have shape (n, 2, {self.nop})')
        ref = math.nan if ref_energy is None else float(ref_energy)
        rc = self._L.qmcb_dmc_init(self._h, C.byref(params), ptr(ini_confs),
                                   ini_confs.shape[0], ref,
                                   global_slot_offset)
        self._check(rc, 'qmcb_dmc_init')
        self._dmc_params = params
        self._dmc_cap = int(params.local_capacity or params.max_num_walkers)

    def dmc_set_state(self, params: DMCParams, confs, energy, weight,
                      scalars: StateScalars, slot_energy=None,
                      global_slot_offset=0):
        confs, energy, weight = _f64(confs), _f64(energy), _f64(weight)
        if slot_energy is not None:
            slot_energy = _f64(slot_energy)
        rc = self._L.qmcb_dmc_set_state(
            self._h, C.byref(params), ptr(confs), ptr(energy), ptr(weight),
            ptr(slot_energy), C.byref(scalars), global_slot_offset)
        self._check(rc, 'qmcb_dmc_set_state')
        self._dmc_params = params
        self._dmc_cap = int(params.local_capacity or params.max_num_walkers)

    def dmc_run_block(self, nts, eval_estimators=False, out=None,
                      density=None, ssf=None):
        """Advance ``nts`` time steps; returns the per-step series (dict of
        arrays of length nts).  ``out`` may hold preallocated arrays."""
        if out is None:
            out = dict(energy=np.zeros(nts), weight=np.zeros(nts),
                       num_walkers=np.zeros(nts, dtype=np.uint64),
                       ref_energy=np.zeros(nts), accum_energy=np.zeros(nts))
        rc = self._L.qmcb_dmc_run_block(
            self._h, nts, int(bool(eval_estimators)), ptr(out['energy']),
            ptr(out['weight']), ptr(out['num_walkers']),
            ptr(out['ref_energy']), ptr(out['accum_energy']), ptr(density),
            ptr(ssf))
        self._check(rc, 'qmcb_dmc_run_block')
        return out

    def dmc_advance(self, nts):
        """Advance without copying any series back (benchmark helper)."""
        rc = self._L.qmcb_dmc_run_block(self._h, nts, 0, None, None, None,
                                        None, None, None, None)
        self._check(rc, 'qmcb_dmc_run_block')

    def dmc_get_state(self, want_confs=True):
        cap, n = self._dmc_cap, self.nop
        out = dict(confs=np.empty((cap, 2, n)) if want_confs else None,
                   energy=np.empty(cap), weight=np.empty(cap),
                   mask=np.empty(cap, dtype=np.uint8),
                   cloning_ref=np.empty(cap, dtype=np.int64))
        sc = StateScalars()
        rc = self._L.qmcb_dmc_get_state(
            self._h, ptr(out['confs']), ptr(out['energy']),
            ptr(out['weight']), ptr(out['mask']), ptr(out['cloning_ref']),
            C.byref(sc))
        self._check(rc, 'qmcb_dmc_get_state')
        out['scalars'] = sc
        return out

    def dmc_scalars(self):
        sc = StateScalars()
        rc = self._L.qmcb_dmc_get_state(self._h, None, None, None, None,
                                        None, C.byref(sc))
        self._check(rc, 'qmcb_dmc_get_state')
        return sc

    def dmc_get_next(self):
        sc = StateScalars()
        rc = self._L.qmcb_dmc_get_next(self._h, None, None, None, None,
                                       C.byref(sc))
        self._check(rc, 'qmcb_dmc_get_next')
        n, cap = int(sc.num_walkers), self._dmc_cap
        out = dict(confs=np.empty((n, 2, self.nop)), energy=np.empty(n),
                   weight=np.empty(n), slot_energy=np.empty(cap))
        rc = self._L.qmcb_dmc_get_next(
            self._h, ptr(out['confs']), ptr(out['energy']),
            ptr(out['weight']), ptr(out['slot_energy']), C.byref(sc))
        self._check(rc, 'qmcb_dmc_get_next')
        out['scalars'] = sc
        return out

    def dmc_get_next_into(self, confs, energy, weight, slot_energy):
        """Copy the evolved population into caller-owned (ideally pinned)
        arrays sized for the capacity; returns the scalars."""
        sc = StateScalars()
        rc = self._L.qmcb_dmc_get_next(
            self._h, ptr(confs), ptr(energy), ptr(weight), ptr(slot_energy),
            C.byref(sc))
        self._check(rc, 'qmcb_dmc_get_next')
        return sc

    # -- multi-GPU ----------------------------------------------------------
    def comm_init(self, unique_id: bytes, world_size: int, rank: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        rc = self._L.qmcb_comm_init(self._h, buf, world_size, rank)
        self._check(rc, 'qmcb_comm_init')

    def comm_init_torch(self, dist, rank, world_size):
        """Create the NCCL communicator, shipping the unique id through an
        initialised ``torch.distributed`` process group (plumbing only)."""
        uid = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        self.comm_init(uid[0], world_size, rank)

    def dmc_rebalance(self):
        moved = C.c_int64()
        rc = self._L.qmcb_dmc_rebalance(self._h, C.byref(moved))
        self._check(rc, 'qmcb_dmc_rebalance')
        return moved.value

    #: dtype of the reference's reblocking tables (stats/reblock.py:436-441)
    OTF_DTYPE = np.dtype([('BLOCK_SIZE', np.int64), ('MEANS', np.float64),
                          ('MEANS_SQR', np.float64),
                          ('NUM_BLOCKS', np.int64)])
    REBLOCK_SERIES = ('energy', 'weight', 'num_walkers', 'ref_energy',
                      'accum_energy')

    def dmc_reblock_reset(self, max_order):
        """Start (``max_order`` >= 0) or stop (None / < 0) the on-device
        reblocking accumulators of the per-step series."""
        mo = -1 if max_order is None else int(max_order)
        rc = self._L.qmcb_dmc_reblock_reset(self._h, mo)
        self._check(rc, 'qmcb_dmc_reblock_reset')
        self._rb_orders = mo + 1

    def dmc_reblock_get(self):
        """dict series name -> structured array (orders,) with the fields of
        the reference's ``otf_data_dtype``: feed it to
        ``stats.reblock.OTFObject`` as is."""
        K = getattr(self, '_rb_orders', 0)
        sums, sqr = np.zeros((5, K)), np.zeros((5, K))
        nblk = np.zeros((5, K), dtype=np.int64)
        rc = self._L.qmcb_dmc_reblock_get(self._h, ptr(sums), ptr(sqr),
                                          ptr(nblk))
        self._check(rc, 'qmcb_dmc_reblock_get')
        out = {}
        for c, name in enumerate(self.REBLOCK_SERIES):
            t_ = np.zeros(K, dtype=self.OTF_DTYPE)
            t_['BLOCK_SIZE'] = 1 << np.arange(K)
            t_['MEANS'], t_['MEANS_SQR'] = sums[c], sqr[c]
            t_['NUM_BLOCKS'] = nblk[c]
            out[name] = t_
        return out

    def set_profiling(self, on=True):
        self._check(self._L.qmcb_set_profiling(self._h, int(on)),
                    'qmcb_set_profiling')

    def last_block_stats(self):
        t, s, n = C.c_double(), C.c_double(), C.c_int64()
        rc = self._L.qmcb_last_block_stats(self._h, C.byref(t), C.byref(s),
                                           C.byref(n))
        self._check(rc, 'qmcb_last_block_stats')
        return dict(total_ms=t.value, step_kernel_ms=s.value,
                    launches=n.value)

    # -- VMC ----------------------------------------------------------------
    def vmc_init(self, confs, move_spread, rng_seed, lower_bound,
                 upper_bound, ssf_num_modes=0, chain_offset=0, proposal=0):
        confs = _f64(confs)
        if confs.ndim == 2:
            confs = confs[None]
        p = VMCParams()
        p.move_spread = move_spread
        p.lower_bound, p.upper_bound = lower_bound, upper_bound
        p.rng_seed = int(rng_seed) & 0xFFFFFFFFFFFFFFFF
        p.chain_offset = chain_offset
        p.ssf_num_modes = ssf_num_modes
        p.proposal = proposal
        rc = self._L.qmcb_vmc_init(self._h, C.byref(p), ptr(confs),
                                   confs.shape[0])
        self._check(rc, 'qmcb_vmc_init')
        self._vmc_chains = confs.shape[0]
        self._vmc_modes = ssf_num_modes

    def vmc_run_block(self, ns, series=True, sums=False, out=None):
        """One block of ``ns`` steps.  ``out`` may hold preallocated result
        arrays by name (e.g. page-locked ones from `pinned_empty`, for a
        full-rate copy back); the others are allocated here."""
        c, m = self._vmc_chains, self._vmc_modes
        given = out or {}

        def buf(name, shape, dtype=np.float64):
            a = given.get(name)
            if a is None:
                return np.empty(shape, dtype=dtype)
            if (a.shape != tuple(shape) or a.dtype != np.dtype(dtype)
                    or not a.flags.c_contiguous or not a.flags.writeable):
                raise ValueError(f'out[{name!r}] must be a writeable '
                                 f'C-contiguous {np.dtype(dtype)} array of '
                                 f'shape {tuple(shape)}')
            return a

        out = {}
        if series:
            out.update(lnpsi=buf('lnpsi', (c, ns)),
                       energy=buf('energy', (c, ns)),
                       move_stat=buf('move_stat', (c, ns), np.uint8),
                       ssf=buf('ssf', (c, ns, m, 3)) if m else None)
        out['accept_rate'] = buf('accept_rate', (c,))
        if sums:
            out['sum_energy'] = buf('sum_energy', (c, 2))
            out['sum_ssf'] = buf('sum_ssf', (c, m, 3)) if m else None
        rc = self._L.qmcb_vmc_run_block(
            self._h, ns, ptr(out.get('lnpsi')), ptr(out.get('energy')),
            ptr(out.get('move_stat')), ptr(out.get('ssf')),
            ptr(out['accept_rate']), ptr(out.get('sum_energy')),
            ptr(out.get('sum_ssf')))
        self._check(rc, 'qmcb_vmc_run_block')
        return out

    def vmc_run_chain(self, ns):
        """``ns`` steps with every state kept: dict(confs (C, ns, 2, N),
        lnpsi, energy, move_stat (C, ns), accept_rate (C,)), one launch."""
        c = self._vmc_chains
        out = dict(confs=np.empty((c, ns, 2, self.nop)),
                   lnpsi=np.empty((c, ns)), energy=np.empty((c, ns)),
                   move_stat=np.empty((c, ns), dtype=np.uint8),
                   accept_rate=np.empty(c))
        rc = self._L.qmcb_vmc_run_chain(
            self._h, ns, ptr(out['lnpsi']), ptr(out['energy']),
            ptr(out['move_stat']), ptr(out['confs']), ptr(out['accept_rate']))
        self._check(rc, 'qmcb_vmc_run_chain')
        return out

    def vmc_one_body_density(self, offsets):
        """g1 of the current state of every chain at ``offsets`` (S,) ->
        (C, S), evaluated on the device."""
        offsets = _f64(np.atleast_1d(offsets))
        out = np.empty((self._vmc_chains, offsets.shape[0]))
        rc = self._L.qmcb_vmc_one_body_density(
            self._h, ptr(offsets), offsets.shape[0], ptr(out))
        self._check(rc, 'qmcb_vmc_one_body_density')
        return out

    def vmc_get_state(self, out=None):
        """(confs (C, 2, N), ln|Psi| (C,)) of the current states; ``out`` may
        be a preallocated (confs, lnpsi) pair to fill."""
        c = self._vmc_chains
        if out is not None:
            confs, ln = out
            for a, shape in ((confs, (c, 2, self.nop)), (ln, (c,))):
                if (a.shape != shape or a.dtype != np.float64
                        or not a.flags.c_contiguous):
                    raise ValueError('out must hold C-contiguous float64 '
                                     f'arrays of shape {(c, 2, self.nop)} '
                                     f'and {(c,)}')
        else:
            confs, ln = np.empty((c, 2, self.nop)), np.empty(c)
        rc = self._L.qmcb_vmc_get_state(self._h, ptr(confs), ptr(ln))
        self._check(rc, 'qmcb_vmc_get_state')
        return confs, ln
