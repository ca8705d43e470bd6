"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of ``libqmc_oracle.so``.

The oracle is the CPU restatement of the reference's hot path
(``oracle/qmc_oracle.c``).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libqmc_oracle.so')

NPARAMS = 25

_f64p = np.ctypeslib.ndpointer(np.float64, flags='C_CONTIGUOUS')
_i64p = np.ctypeslib.ndpointer(np.int64, flags='C_CONTIGUOUS')
_u64p = np.ctypeslib.ndpointer(np.uint64, flags='C_CONTIGUOUS')
_u8p = np.ctypeslib.ndpointer(np.uint8, flags='C_CONTIGUOUS')


def build(force=False):
    src = os.path.join(_HERE, 'qmc_oracle.c')
    if (force or not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)):
        subprocess.check_call(['make', '-C', _HERE, '-B', 'libqmc_oracle.so'],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


class StateBuf(C.Structure):
    _fields_ = [('confs', C.c_void_p), ('energy', C.c_void_p),
                ('weight', C.c_void_p), ('mask', C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    opt = C.c_void_p
    L.qmco_model_eval.argtypes = [_f64p, _f64p, C.c_int64, opt, opt, opt]
    L.qmco_model_eval.restype = None
    L.qmco_fourier_density.argtypes = [_f64p, _f64p, C.c_int64, C.c_int, _f64p]
    L.qmco_fourier_density.restype = None
    L.qmco_fourier_density_k.argtypes = [_f64p, _f64p, C.c_int64, _f64p,
                                         C.c_int64, _f64p]
    L.qmco_fourier_density_k.restype = None
    L.qmco_one_body_density.argtypes = [_f64p, _f64p, C.c_int64, _f64p,
                                        C.c_int64, _f64p]
    L.qmco_one_body_density.restype = None
    L.qmco_rng_uniform2.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32,
                                    C.c_uint32, C.c_uint32, _f64p]
    L.qmco_rng_normal4.argtypes = L.qmco_rng_uniform2.argtypes
    L.qmco_rng_uniform4.argtypes = L.qmco_rng_uniform2.argtypes
    L.qmco_branch.argtypes = [_f64p, C.c_int64, C.c_int64, _f64p, _i64p]
    L.qmco_branch.restype = C.c_int64
    L.qmco_evolve_state.argtypes = [
        _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, _u8p, _f64p, _f64p, _f64p,
        C.c_int64, C.c_int64, C.c_double, C.c_double, _i64p, _f64p,
        C.c_double, C.c_double, C.c_int]
    L.qmco_evolve_state.restype = None
    L.qmco_prepare_state.argtypes = [_f64p, _f64p, C.c_int64, C.c_int64,
                                     _f64p, _f64p, _f64p, _u8p]
    L.qmco_prepare_state.restype = None
    L.qmco_ssf_step.argtypes = [_f64p, C.c_int64, _f64p, C.c_int64, C.c_int64,
                                _i64p, C.c_int, C.c_int, C.c_int64, _f64p,
                                _f64p]
    L.qmco_ssf_step.restype = None
    L.qmco_density_step.argtypes = [_f64p, C.c_int64, _f64p, C.c_int64,
                                    C.c_int64, C.c_int, C.c_int, C.c_int64,
                                    _f64p, _f64p]
    L.qmco_density_step.restype = None
    L.qmco_dmc_block.argtypes = [
        _f64p, C.c_uint64, C.c_double, C.c_int64, C.c_double, C.c_int64,
        C.c_double, C.c_double, C.c_int,
        C.POINTER(StateBuf), C.POINTER(StateBuf), C.POINTER(StateBuf),
        _i64p, _f64p, _i64p, C.c_int64,
        _f64p, _f64p, _u64p, _f64p, _f64p,
        C.c_int, C.c_int, C.c_int, C.c_int64, opt, opt,
        C.c_int, C.c_int, C.c_int64, opt, opt, opt, opt]
    L.qmco_dmc_block.restype = None
    L.qmco_vmc_block.argtypes = [
        _f64p, C.c_uint64, C.c_double, C.c_double, C.c_double,
        C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int,
        _f64p, _f64p, _f64p, opt, C.c_int,
        _f64p, _f64p, _u8p, opt, _f64p, opt, C.c_int]
    L.qmco_vmc_block.restype = None
    L.qmco_num_threads.restype = C.c_int
    L.qmco_set_num_threads.argtypes = [C.c_int]
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _params(p):
    p = np.ascontiguousarray(p, dtype=np.float64)
    assert p.shape == (NPARAMS,)
    return p


def num_threads():
    return lib().qmco_num_threads()


def set_num_threads(n):
    lib().qmco_set_num_threads(int(n))


def model_eval(params, confs, want=('lnpsi', 'energy', 'drift')):
    """confs [B,2,N] -> dict(lnpsi[B], energy[B], drift[B,N])."""
    p = _params(params)
    confs = np.ascontiguousarray(confs, dtype=np.float64)
    if confs.ndim == 2:
        confs = confs[None]
    nconf, _, nop = confs.shape
    assert nop == int(p[3])
    out = {}
    out['lnpsi'] = np.empty(nconf) if 'lnpsi' in want else None
    out['energy'] = np.empty(nconf) if 'energy' in want else None
    out['drift'] = np.empty((nconf, nop)) if 'drift' in want else None
    lib().qmco_model_eval(p, confs, nconf, _ptr(out['lnpsi']),
                          _ptr(out['energy']), _ptr(out['drift']))
    return out


def fourier_density(params, confs, num_modes):
    p = _params(params)
    confs = np.ascontiguousarray(confs, dtype=np.float64)
    if confs.ndim == 2:
        confs = confs[None]
    out = np.empty((confs.shape[0], num_modes, 3))
    lib().qmco_fourier_density(p, confs, confs.shape[0], num_modes, out)
    return out


def fourier_density_k(params, confs, kz_set):
    """confs [B,2,N], kz_set [K] -> complex [B,K]."""
    p = _params(params)
    confs = np.ascontiguousarray(confs, dtype=np.float64)
    if confs.ndim == 2:
        confs = confs[None]
    kz = np.ascontiguousarray(kz_set, dtype=np.float64)
    out = np.empty((confs.shape[0], len(kz), 2))
    lib().qmco_fourier_density_k(p, confs, confs.shape[0], kz, len(kz), out)
    return out[..., 0] + 1j * out[..., 1]


def weighed_variance(wf_abs_log_set, ini_wf_abs_log_set, energy_set):
    """Correlated-sampling objective: qmc_base/jastrow/model.py:1147-1165
    with the weights of principal_function (:1199-1203).  Returns
    (variance, weighted mean energy)."""
    wl = 2 * (np.asarray(wf_abs_log_set) - np.asarray(ini_wf_abs_log_set))
    energy_set = np.asarray(energy_set)
    rel_weights = np.exp(wl - wl.max())
    weight_sum = rel_weights.sum()
    ref_energy = (rel_weights * energy_set).sum() / weight_sum
    e_diff = rel_weights * (energy_set - ref_energy) ** 2
    return e_diff.sum() / weight_sum, ref_energy


OTF_DTYPE = np.dtype([('BLOCK_SIZE', np.int64), ('MEANS', np.float64),
                      ('MEANS_SQR', np.float64), ('NUM_BLOCKS', np.int64)])


def otf_create(source_data):
    """On-the-fly reblocking tables of a series [n] or [n, ncols]:
    stats/reblock.py:447-457 (order), :508-522 (init), :525-604 (the loop),
    statement by statement.  Returns a structured array [ncols, order + 1]
    ([order + 1] for a 1-d series) with the reference's otf_data_dtype."""
    import math
    src = np.asarray(source_data, dtype=np.float64)
    one_d = src.ndim == 1
    if one_d:
        src = src[:, None]
    n, ncols = src.shape
    max_order = int(math.floor(math.log(n) / math.log(2)))
    otf = np.zeros((ncols, max_order + 1), dtype=OTF_DTYPE)
    for order in range(max_order + 1):
        otf['BLOCK_SIZE'][:, order] = 1 << order
    means = np.zeros((ncols, max_order + 1, 2))
    for index in range(n):
        order, block_size, next_block_size = 0, 1, 2
        mean_index = index % 2
        for nc in range(ncols):
            v = src[index, nc]
            otf['MEANS'][nc, 0] += v
            otf['MEANS_SQR'][nc, 0] += v * v        # numba lowers x ** 2 to x * x
            means[nc, 0, mean_index] = v
            otf['NUM_BLOCKS'][nc, 0] += 1
        while order < max_order and not (index + 1) % next_block_size:
            order += 1
            block_size <<= 1
            block_index = (index + 1) // block_size - 1
            mean_index = block_index % 2
            for nc in range(ncols):
                m = means[nc, order - 1].mean()
                otf['MEANS'][nc, order] += m
                otf['MEANS_SQR'][nc, order] += m * m
                means[nc, order, mean_index] = m
                otf['NUM_BLOCKS'][nc, order] += 1
            next_block_size = block_size << 1
    return otf[0] if one_d else otf


def otf_update(obj_data, ext_obj_data):
    """stats/reblock.py:927-948: accumulate a compatible table in place."""
    assert np.all(obj_data['BLOCK_SIZE'] == ext_obj_data['BLOCK_SIZE'])
    obj_data['MEANS'] += ext_obj_data['MEANS']
    obj_data['MEANS_SQR'] += ext_obj_data['MEANS_SQR']
    obj_data['NUM_BLOCKS'] += ext_obj_data['NUM_BLOCKS']


def one_body_density(params, confs, offsets):
    """confs [B,2,N], offsets [S] -> [B,S]."""
    p = _params(params)
    confs = np.ascontiguousarray(confs, dtype=np.float64)
    if confs.ndim == 2:
        confs = confs[None]
    offsets = np.ascontiguousarray(offsets, dtype=np.float64)
    out = np.empty((confs.shape[0], len(offsets)))
    lib().qmco_one_body_density(p, confs, confs.shape[0], offsets,
                                len(offsets), out)
    return out


def rng_uniform2(seed, c0, c1, c2, stream):
    out = np.empty(2)
    lib().qmco_rng_uniform2(seed, c0, c1, c2, stream, out)
    return out


def rng_normal4(seed, c0, c1, c2, stream):
    out = np.empty(4)
    lib().qmco_rng_normal4(seed, c0, c1, c2, stream, out)
    return out


def rng_uniform4(seed, c0, c1, c2, stream):
    out = np.empty(4)
    lib().qmco_rng_uniform4(seed, c0, c1, c2, stream, out)
    return out


def branch(weights, prev_num_walkers, max_num_walkers, uniforms):
    ref = np.zeros(max_num_walkers, dtype=np.int64)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    u = np.ascontiguousarray(uniforms, dtype=np.float64)
    n = lib().qmco_branch(w, prev_num_walkers, max_num_walkers, u, ref)
    return int(n), ref


class DMCState:
    """The three reference buffers (prev / actual / next) + scalars."""

    def __init__(self, params, ini_confs, max_num_walkers, ref_energy=None):
        p = _params(params)
        self.params = p
        nop = int(p[3])
        ini_confs = np.ascontiguousarray(ini_confs, dtype=np.float64)
        n = ini_confs.shape[0]
        wmax = int(max_num_walkers)
        self.wmax, self.nop = wmax, nop
        confs = np.zeros((wmax, 2, nop))
        energy = np.zeros(wmax)
        weight = np.zeros(wmax)
        mask = np.zeros(wmax, dtype=np.uint8)
        lib().qmco_prepare_state(p, ini_confs, n, wmax, confs, energy,
                                 weight, mask)
        # mrbp_qmc/dmc.py:299-312
        state_energy = float((energy[:n] * weight[:n]).sum())
        state_weight = float(weight[:n].sum())
        self.ini_energy = state_energy
        self.ini_weight = state_weight
        if ref_energy is None:
            ref_energy = state_energy / state_weight
        # qmc_base/dmc.py:707-716: three copies of the initial state
        self.bufs = [dict(confs=confs.copy(), energy=energy.copy(),
                          weight=weight.copy(), mask=mask.copy())
                     for _ in range(3)]   # prev, act, next
        self.ref = np.zeros(wmax, dtype=np.int64)
        self.scal = np.array([ref_energy, 0., 0.])
        self.cnt = np.array([n, 0], dtype=np.int64)

    @property
    def prev(self):
        return self.bufs[0]

    @property
    def act(self):
        return self.bufs[1]

    @property
    def next(self):
        return self.bufs[2]

    @property
    def num_walkers(self):
        return int(self.cnt[0])

    def _sb(self, d):
        return StateBuf(d['confs'].ctypes.data, d['energy'].ctypes.data,
                        d['weight'].ctypes.data, d['mask'].ctypes.data)

    def evolve(self, num_walkers, time_step, ref_energy, normals,
               z_min, z_max, energy_mode=0):
        """One evolve_state call with explicit normals [Wmax, N]."""
        normals = np.ascontiguousarray(normals, dtype=np.float64)
        pr, ac, nx = self.bufs
        lib().qmco_evolve_state(
            self.params, pr['confs'], pr['energy'], ac['confs'],
            ac['energy'], ac['weight'], ac['mask'], nx['confs'],
            nx['energy'], nx['weight'], num_walkers, self.wmax, time_step,
            ref_energy, self.ref, normals, z_min, z_max, energy_mode)

    def swap(self):
        self.bufs[0], self.bufs[2] = self.bufs[2], self.bufs[0]

    def run_block(self, seed, time_step, target_num_walkers, nwc_factor,
                  nts, z_min, z_max, energy_mode=0, eval_est=False,
                  ssf=None, density=None, uniforms_ext=None,
                  normals_ext=None):
        """ssf / density: dict(num, pure, pfw, iter, aux) or None.
        uniforms_ext [nts, Wmax], normals_ext [nts, Wmax, N] (unit variance)
        replace the Philox streams when given."""
        if uniforms_ext is not None:
            uniforms_ext = np.ascontiguousarray(uniforms_ext, np.float64)
            assert uniforms_ext.shape == (nts, self.wmax)
        if normals_ext is not None:
            normals_ext = np.ascontiguousarray(normals_ext, np.float64)
            assert normals_ext.shape == (nts, self.wmax, self.nop)
        it = dict(energy=np.zeros(nts), weight=np.zeros(nts),
                  num_walkers=np.zeros(nts, dtype=np.uint64),
                  ref_energy=np.zeros(nts), accum_energy=np.zeros(nts))
        sbs = [self._sb(b) for b in self.bufs]
        sm, sp, sw, si, sa = (0, 0, 0, None, None)
        if ssf is not None:
            sm, sp, sw = ssf['num'], int(ssf['pure']), ssf['pfw']
            si, sa = ssf['iter'], ssf['aux']
        dm, dp, dw, di, da = (0, 0, 0, None, None)
        if density is not None:
            dm, dp, dw = density['num'], int(density['pure']), density['pfw']
            di, da = density['iter'], density['aux']
        lib().qmco_dmc_block(
            self.params, seed, time_step, target_num_walkers, nwc_factor,
            self.wmax, z_min, z_max, energy_mode,
            C.byref(sbs[0]), C.byref(sbs[1]), C.byref(sbs[2]),
            self.ref, self.scal, self.cnt, nts,
            it['energy'], it['weight'], it['num_walkers'], it['ref_energy'],
            it['accum_energy'], int(eval_est),
            sm, sp, sw, _ptr(si), _ptr(sa), dm, dp, dw, _ptr(di), _ptr(da),
            _ptr(uniforms_ext), _ptr(normals_ext))
        # the C side swapped struct contents if roles changed: re-map
        ptr2buf = {b['confs'].ctypes.data: b for b in self.bufs}
        self.bufs = [ptr2buf[sb.confs] for sb in sbs]
        return it


def ssf_step(params, step_idx, confs, num_walkers, wmax, cloning_ref,
             num_modes, pure, pfw, iter_ssf, aux):
    lib().qmco_ssf_step(_params(params), step_idx, confs, num_walkers, wmax,
                        cloning_ref, num_modes, int(pure), pfw, iter_ssf, aux)


def density_step(params, step_idx, confs, num_walkers, wmax, num_bins, pure,
                 pfw, iter_density, aux):
    lib().qmco_density_step(_params(params), step_idx, confs, num_walkers,
                            wmax, num_bins, int(pure), pfw, iter_density, aux)


def vmc_block(params, seed, move_spread, z_min, z_max, cur, lnpsi_cur,
              energy_prev, ssf_prev, num_modes, ns, step0, first,
              chain_offset=0, uniforms_ext=None, proposal=0):
    """Advance C chains by ns yielded states. cur [C,2,N] is updated.
    proposal: 0 uniform (move_spread = width), 1 gaussian (= sigma)."""
    p = _params(params)
    nch = cur.shape[0]
    out = dict(lnpsi=np.zeros((nch, ns)), energy=np.zeros((nch, ns)),
               stat=np.zeros((nch, ns), dtype=np.uint8),
               ssf=(np.zeros((nch, ns, num_modes, 3)) if num_modes else None),
               accept_rate=np.zeros(nch))
    if uniforms_ext is not None:
        uniforms_ext = np.ascontiguousarray(uniforms_ext, dtype=np.float64)
    lib().qmco_vmc_block(p, seed, move_spread, z_min, z_max, nch,
                         chain_offset, ns, step0, int(first), cur, lnpsi_cur,
                         energy_prev, _ptr(ssf_prev), num_modes,
                         out['lnpsi'], out['energy'], out['stat'],
                         _ptr(out['ssf']), out['accept_rate'],
                         _ptr(uniforms_ext), int(proposal))
    return out
