"""TEST INFRASTRUCTURE ONLY -- freeze outputs of the LIVE reference.

Runs the unmodified reference (``/root/reference/src``, through
``oracle/refshim.py``) in the builder container and writes small ``.npz``
fixtures to ``tests/golden/``.  The reference's own tests hold no golden
vectors for this path, so these files are the pin for
both the C oracle and the CUDA engine.

    python oracle/make_golden.py            # the deterministic fixtures and
                                            # the N=16/20 statistical runs
    python oracle/make_golden.py statpure statpure100 statpure200 statvmc50
                                            # the long runs at the BASELINE
                                            # particle numbers (N = 50, 100,
                                            # 200 DMC with the pure estimators,
                                            # N = 50 VMC): minutes each

Every random input is drawn from a seeded numpy Generator; every random
number the reference consumes *inside* its JIT code is re-drawn afterwards
from Numba's own generator with the same seed and call sequence, so the
fixtures also contain the exact uniforms / gaussians the reference used.
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
PI = math.pi

from specs import SPECS  # noqa: E402


def param_block(spec):
    flat = [float(v) for part in (spec.params, spec.obf_params,
                                  spec.tbf_params) for v in tuple(part)]
    return np.array(flat, dtype=np.float64)


def make_confs(rng, spec, nconf):
    nop, size = spec.boson_number, spec.supercell_size
    rm = abs(spec.tbf_contact_cutoff)
    confs = np.zeros((nconf, 2, nop))
    confs[:, 0, :] = rng.random((nconf, nop)) * size
    # Row 1 (drift) is ignored by the model functions: fill with junk.
    confs[:, 1, :] = rng.standard_normal((nconf, nop))
    # Edge cases.
    confs[0, 0, :] = np.linspace(0, size, nop, endpoint=False)   # regular
    if nconf > 1 and nop >= 4:
        c = confs[1, 0]
        c[0] = 0.0                              # lower boundary
        c[1] = size * (1 - 2.0 ** -30)          # close pair across the wrap
        c[2] = c[3] + 1e-9                      # nearly coincident pair
    if nconf > 2 and nop >= 6:
        c = confs[2, 0]
        c[1] = (c[0] + rm) % size               # pair exactly at r_m
        c[3] = (c[2] + 0.5 * size) % size       # pair at L/2
        c[5] = size                             # recast quirk Q5: z == L
    if nconf > 3 and nop >= 4:
        c = confs[3, 0]
        wa = spec.well_width
        c[0] = 3 + wa                           # exactly on the well edge
        c[1] = 3 + wa + 1e-9                    # just inside the barrier
        c[2] = 2.0                              # cell boundary
    return confs


def gen_model(mrbp, name, kwargs, rng):
    model = mrbp.model
    spec = model.Spec(**kwargs)
    cf = model.core_funcs
    nop = spec.boson_number
    nconf = 4 if nop >= 200 else 12
    confs = make_confs(rng, spec, nconf)
    cfc = spec.cfc_spec
    lnpsi = np.array([cf.wf_abs_log(c, *cfc) for c in confs])
    energy = np.array([cf.energy(c, *cfc) for c in confs])
    drift = np.array([cf.drift(c, *cfc)[1] for c in confs])
    e_and_d = np.array([[cf.ith_energy_and_drift(i, c, *cfc)
                         for i in range(nop)] for c in confs])
    num_modes = min(2 * nop, 24)
    momenta = np.arange(num_modes) * 2 * PI / spec.supercell_size
    fdk = np.array([[cf.fourier_density(k, c, *cfc) for k in momenta]
                    for c in confs])
    ssf = np.stack([(fdk * fdk.conj()).real, fdk.real, fdk.imag], axis=-1)
    # one-body density matrix at a few displacements, incl. beyond the box
    size = spec.supercell_size
    obd_offsets = np.array([0., 0.05, 0.37, 1.0, 2.5, 0.5 * size, -0.3,
                            -1.7 * size])
    nobd = min(len(confs), 4)
    obd = np.array([[cf.one_body_density(sz, c, *cfc) for sz in obd_offsets]
                    for c in confs[:nobd]])
    # rho_k at momenta that are NOT multiples of 2 pi / L (the gufunc of
    # PhysicalFuncs takes any kz_set)
    kz_set = np.array([0.0, 0.3, 1.234, -2.5, 7.0, 2 * PI / size])
    fdk_k = np.array([[cf.fourier_density(k, c, *cfc) for k in kz_set]
                      for c in confs])
    out = dict(spec_keys=np.array(list(kwargs.keys())),
               spec_vals=np.array([float(v) for v in kwargs.values()]),
               kz_set=kz_set, fdk_k=fdk_k,
               params=param_block(spec), confs=confs, lnpsi=lnpsi,
               energy=energy, drift=drift, ith_energy=e_and_d[..., 0],
               ssf=ssf, num_modes=num_modes, obd_offsets=obd_offsets,
               obd=obd)
    np.savez_compressed(os.path.join(GOLDEN, f'model_{name}.npz'), **out)
    print(f'model_{name}: lnpsi[0]={lnpsi[0]:.15g} E[0]={energy[0]:.15g}')


def numba_draws():
    import numba as nb

    @nb.njit
    def draw_uniform(seed, n):
        np.random.seed(seed)
        out = np.empty(n)
        for i in range(n):
            out[i] = np.random.rand()
        return out

    @nb.njit
    def draw_dmc(seed, counts_u, counts_n, nop):
        """Replay the serial DMC call sequence: per step, counts_u[t] calls
        of rand() (branching), then counts_n[t]*nop calls of normal()."""
        np.random.seed(seed)
        nts = counts_u.shape[0]
        wmax_u = counts_u.max()
        wmax_n = counts_n.max()
        uni = np.zeros((nts, wmax_u))
        nor = np.zeros((nts, wmax_n, nop))
        for t in range(nts):
            for s in range(counts_u[t]):
                uni[t, s] = np.random.rand()
            for s in range(counts_n[t]):
                for i in range(nop):
                    nor[t, s, i] = np.random.normal(0., 1.)
        return uni, nor

    @nb.njit
    def draw_vmc_ndf(seed, total, nop):
        """Replay the gaussian-proposal VMC call sequence: per step nop calls
        of normal(0, 1) then one rand()."""
        np.random.seed(seed)
        out = np.empty((total, nop + 1))
        for t in range(total):
            for i in range(nop):
                out[t, i] = np.random.normal(0., 1.)
            out[t, nop] = np.random.rand()
        return out

    draw_uniform.ndf = draw_vmc_ndf
    return draw_uniform, draw_dmc


def gen_dmc_step(mrbp, name, kwargs, rng):
    """One bit-reproducible evolve_state call (sigma = 0), hand-made cloning
    table with deaths and duplicates, poisoned `actual` energies (quirk Q1).
    """
    model, dmc = mrbp.model, mrbp.dmc
    spec = model.Spec(**kwargs)
    nop = spec.boson_number
    wmax, n_ini = 24, 16
    dt = 1e-3
    sampling = dmc.Sampling(spec, dt, wmax, n_ini, rng_seed=1,
                            jit_parallel=False)
    cf = sampling.core_funcs
    confs = np.zeros((n_ini, 2, nop))
    confs[:, 0, :] = rng.random((n_ini, nop)) * spec.supercell_size
    ini_state = sampling.build_state(confs)
    cfc = sampling.cfc_spec
    prev = cf.init_state_data_from_state(ini_state, cfc)
    act = cf.init_state_data_from_state(ini_state, cfc)
    nxt = cf.init_state_data_from_state(ini_state, cfc)
    # Poison the persistent per-slot energies so Q1 is visible.
    act.props.energy[:] = ini_state.props.energy + \
        rng.standard_normal(wmax) * 3.0
    act_energy_in = act.props.energy.copy()
    bs = cf.init_branching_spec(wmax)
    ref = np.array([0, 0, 1, 3, 3, 3, 4, 6, 7, 7, 8, 10, 11, 12, 13, 14, 15,
                    15, 15, 2], dtype=np.int64)
    nw = len(ref)
    bs.cloning_ref[:nw] = ref
    ref_energy = float(ini_state.ref_energy) + 0.37
    cfc0 = cfc._replace(ddf_params=cfc.ddf_params._replace(sigma_spread=0.0))
    cf.evolve_state(prev, act, nxt, nw, wmax, dt, ref_energy, bs, cfc0)
    z_min, z_max = spec.boundaries
    out = dict(params=param_block(spec), ini_confs=confs,
               ini_state_confs=np.asarray(ini_state.confs),
               ini_state_energy=np.asarray(ini_state.props.energy),
               ini_state_weight=np.asarray(ini_state.props.weight),
               ini_state_mask=np.asarray(ini_state.props.mask),
               ini_ref_energy=float(ini_state.ref_energy),
               ini_energy_sum=float(ini_state.energy),
               act_energy_in=act_energy_in, cloning_ref=bs.cloning_ref.copy(),
               num_walkers=nw, max_num_walkers=wmax, time_step=dt,
               ref_energy=ref_energy, z_min=z_min, z_max=z_max,
               act_confs=act.confs, act_energy=act.props.energy,
               act_weight=act.props.weight, act_mask=act.props.mask,
               next_confs=nxt.confs, next_energy=nxt.props.energy,
               next_weight=nxt.props.weight)
    np.savez_compressed(os.path.join(GOLDEN, f'dmc_step_{name}.npz'), **out)
    print(f'dmc_step_{name}: next_w[:3]={nxt.props.weight[:3]}')


def gen_branch(mrbp, rng, draw_uniform):
    model, dmc = mrbp.model, mrbp.dmc
    spec = model.Spec(**SPECS['ll_n16'])
    cases = {}
    for tag, (wprev, wmax, scale) in dict(
            plain=(40, 64, 1.0), capped=(40, 44, 1.6),
            dying=(40, 64, 0.3)).items():
        sampling = dmc.Sampling(spec, 1e-3, wmax, wprev, rng_seed=1,
                                jit_parallel=False)
        cf = sampling.core_funcs
        sd = cf.init_state_data((wmax,), sampling.cfc_spec)
        w = np.exp(rng.standard_normal(wprev) * 0.5) * scale
        sd.props.weight[:wprev] = w
        bs = cf.init_branching_spec(wmax)
        seed = 1234 + len(cases)
        from phd_qmclib.qmc_base.utils import numba_seed
        numba_seed(seed)
        nw = cf.sync_branching_spec(sd, wprev, wmax, bs)
        uni = draw_uniform(seed, wprev)
        cases[tag] = dict(weights=w, uniforms=uni, wmax=wmax,
                          num_walkers=nw, cloning_ref=bs.cloning_ref.copy())
        print(f'branch_{tag}: W={nw}')
    flat = {f'{t}_{k}': v for t, d in cases.items() for k, v in d.items()}
    np.savez_compressed(os.path.join(GOLDEN, 'dmc_branch.npz'), **flat)


def gen_dmc_blocks(mrbp, name, kwargs, rng, draw_dmc, *, n_ini, wmax, dt,
                   nts, nblocks, num_modes, num_bins, pure, seed,
                   nwc=0.125):
    """The reference's own blocks() in serial mode with estimators, plus the
    exact random numbers it consumed (replayed from Numba's generator)."""
    model, dmc = mrbp.model, mrbp.dmc
    spec = model.Spec(**kwargs)
    nop = spec.boson_number
    ssf_spec = dmc.SSFEstSpec(num_modes, as_pure_est=pure,
                              pfw_num_time_steps=nts)
    den_spec = dmc.DensityEstSpec(num_bins, as_pure_est=pure,
                                  pfw_num_time_steps=nts)
    sampling = dmc.Sampling(spec, dt, wmax, n_ini,
                            num_walkers_control_factor=nwc, rng_seed=seed,
                            density_est_spec=den_spec, ssf_est_spec=ssf_spec,
                            jit_parallel=False)
    confs = np.zeros((n_ini, 2, nop))
    confs[:, 0, :] = rng.random((n_ini, nop)) * spec.supercell_size
    ini_state = sampling.build_state(confs)
    blocks = sampling.blocks(ini_state, nts, 0)
    rec = dict(energy=[], weight=[], num_walkers=[], ref_energy=[],
               accum_energy=[], density=[], ssf=[])
    last = None
    for b, block in zip(range(nblocks), blocks):
        ip = block.iter_props
        rec['energy'].append(ip.energy.copy())
        rec['weight'].append(ip.weight.copy())
        rec['num_walkers'].append(ip.num_walkers.copy())
        rec['ref_energy'].append(ip.ref_energy.copy())
        rec['accum_energy'].append(ip.accum_energy.copy())
        rec['density'].append(block.iter_density.copy())
        rec['ssf'].append(block.iter_ssf.copy())
        last = block.last_state
    nw = np.concatenate(rec['num_walkers']).astype(np.int64)
    assert nw.max() < wmax, 'capacity hit: replay of the RNG would be wrong'
    counts_u = np.concatenate([[n_ini], nw[:-1]]).astype(np.int64)
    uni, nor = draw_dmc(seed, counts_u, nw, nop)
    total = nts * nblocks
    uniforms = np.zeros((total, wmax))
    normals = np.zeros((total, wmax, nop))
    uniforms[:, :uni.shape[1]] = uni
    normals[:, :nor.shape[1]] = nor
    z_min, z_max = spec.boundaries
    out = dict(params=param_block(spec), ini_confs=confs, n_ini=n_ini,
               max_num_walkers=wmax, time_step=dt, nts=nts, nblocks=nblocks,
               nwc_factor=nwc, target_num_walkers=n_ini, z_min=z_min,
               z_max=z_max, num_modes=num_modes, num_bins=num_bins,
               pure=int(pure), uniforms=uniforms, normals=normals,
               last_confs=np.asarray(last.confs),
               last_energy=np.asarray(last.props.energy),
               last_weight=np.asarray(last.props.weight),
               last_mask=np.asarray(last.props.mask),
               last_cloning_ref=np.asarray(last.branching_spec.cloning_ref),
               last_num_walkers=int(last.num_walkers),
               last_ref_energy=float(last.ref_energy),
               last_accum_energy=float(last.accum_energy))
    for k, v in rec.items():
        out['it_' + k] = np.stack(v)
    np.savez_compressed(os.path.join(GOLDEN, f'dmc_blocks_{name}.npz'), **out)
    print(f'dmc_blocks_{name}: W={nw.tolist()} E_ref[-1]='
          f'{out["it_ref_energy"][-1, -1]:.12g}')


def gen_vmc_blocks(mrbp, name, kwargs, rng, draw_uniform, *, move_spread,
                   ns, nblocks, num_modes, seed, gaussian=False):
    """gaussian: the vmc_ndf sampler (mrbp_qmc/vmc_ndf.py), `move_spread` is
    then sigma = sqrt(time_step)."""
    model, vmc = mrbp.model, mrbp.vmc
    spec = model.Spec(**kwargs)
    nop = spec.boson_number
    if gaussian:
        from phd_qmclib.mrbp_qmc import vmc_ndf
        # attrs field order of the reference class: move_spread (inherited,
        # unused by the gaussian proposal), model_spec, time_step, ...
        sampling = vmc_ndf.Sampling(move_spread=move_spread, model_spec=spec,
                                    time_step=move_spread ** 2, rng_seed=seed,
                                    ssf_est_spec=vmc.SSFEstSpec(num_modes))
    else:
        sampling = vmc.Sampling(spec, move_spread, rng_seed=seed,
                                ssf_est_spec=vmc.SSFEstSpec(num_modes))
    conf = spec.get_sys_conf_buffer()
    conf[0, :] = rng.random(nop) * spec.supercell_size
    ini_conf = conf.copy()
    ini_state = sampling.build_state(conf)
    rec = dict(lnpsi=[], energy=[], stat=[], ssf=[], accept_rate=[])
    last = None
    for b, block in zip(range(nblocks), sampling.blocks(ns, ini_state)):
        ip = block.iter_props
        rec['lnpsi'].append(ip.wf_abs_log.copy())
        rec['energy'].append(ip.energy.copy())
        rec['stat'].append(ip.move_stat.copy())
        rec['ssf'].append(block.iter_ssf.copy())
        rec['accept_rate'].append(block.accept_rate)
        last = block.last_state
    total = ns * nblocks - 1          # first yield consumes no RNG
    if gaussian:
        uni = draw_uniform.ndf(seed, total, nop)
    else:
        uni = draw_uniform(seed, total * (nop + 1)).reshape(total, nop + 1)
    z_min, z_max = spec.boundaries
    out = dict(params=param_block(spec), ini_conf=ini_conf,
               ini_lnpsi=float(ini_state.wf_abs_log), move_spread=move_spread,
               ns=ns, nblocks=nblocks, num_modes=num_modes, z_min=z_min,
               z_max=z_max, uniforms=uni, proposal=int(gaussian),
               last_conf=np.asarray(last.sys_conf),
               last_lnpsi=float(last.wf_abs_log))
    for k, v in rec.items():
        out['it_' + k] = np.stack(v)
    np.savez_compressed(os.path.join(GOLDEN, f'vmc_blocks_{name}.npz'), **out)
    print(f'vmc_blocks_{name}: accept={rec["accept_rate"]}')


def _ref_reblock(series):
    """mean and mean_eff_error of a series from the reference's OWN blocking
    analysis (stats/reblock.py:113-217, 327-420)."""
    import warnings
    from phd_qmclib.stats import reblock
    obj = reblock.Object(np.ascontiguousarray(series, dtype=np.float64))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return float(obj.mean), float(obj.mean_eff_error)


def gen_dmc_stat(mrbp, name, kwargs, *, n_target, wmax, dt, nts, nblocks,
                 burn, seed, nwc=0.125, num_modes=8):
    """A longer serial reference DMC run: per-block sums of the energy and of
    the mixed S(k) estimator for the statistical parity tests, with the
    error bars of the reference's own reblocking."""
    model, dmc = mrbp.model, mrbp.dmc
    spec = model.Spec(**kwargs)
    nop = spec.boson_number
    ssf_spec = dmc.SSFEstSpec(num_modes, as_pure_est=False,
                              pfw_num_time_steps=nts)
    sampling = dmc.Sampling(spec, dt, wmax, n_target,
                            num_walkers_control_factor=nwc, rng_seed=seed,
                            ssf_est_spec=ssf_spec, jit_parallel=False)
    rng = np.random.default_rng(seed)
    confs = np.zeros((n_target, 2, nop))
    confs[:, 0, :] = rng.random((n_target, nop)) * spec.supercell_size
    ini_state = sampling.build_state(confs)
    e_sum, w_sum, s_sum = [], [], []
    for b, block in zip(range(burn + nblocks),
                        sampling.blocks(ini_state, nts, burn)):
        if b < burn:
            continue
        e_sum.append(block.iter_props.energy.sum())
        w_sum.append(block.iter_props.weight.sum())
        s_sum.append(np.asarray(block.iter_ssf).sum(axis=0))
    e_sum, w_sum, s_sum = np.array(e_sum), np.array(w_sum), np.array(s_sum)
    # ratio estimators, linearised per block, through the reference reblock
    e_mean = e_sum.sum() / w_sum.sum()
    _, e_err = _ref_reblock((e_sum - e_mean * w_sum) / w_sum.mean())
    sk_mean = s_sum[:, :, 0].sum(axis=0) / w_sum.sum()
    sk_err = np.array([
        _ref_reblock((s_sum[:, m, 0] - sk_mean[m] * w_sum) / w_sum.mean())[1]
        for m in range(num_modes)])
    out = dict(params=param_block(spec), ini_confs=confs, n_target=n_target,
               max_num_walkers=wmax, time_step=dt, nts=nts, nblocks=nblocks,
               burn=burn, nwc_factor=nwc, block_energy=e_sum,
               block_weight=w_sum, block_ssf=s_sum, num_modes=num_modes,
               ref_energy_mean=e_mean, ref_energy_err=e_err,
               ref_sk_mean=sk_mean, ref_sk_err=sk_err)
    np.savez_compressed(os.path.join(GOLDEN, f'dmc_stat_{name}.npz'), **out)
    epn = e_sum / w_sum / nop
    print(f'dmc_stat_{name}: E/N = {e_mean / nop:.6f} +- {e_err / nop:.6f} '
          f'(reference reblock; naive {epn.std(ddof=1) / math.sqrt(len(epn)):.6f})'
          f' <|rho_k|^2>/N = {sk_mean / nop}')


def gen_dmc_stat_pure(mrbp, name, kwargs, *, n_target, wmax, dt, nts, nblocks,
                      burn, seed, nwc=0.5, num_modes=8, num_bins=50,
                      parallel=True):
    """A reference DMC run at a BASELINE particle number with the PURE
    (forward-walking) S(k) and density estimators, reduced per block the way
    qmc_exec/dmc/proc.py:304-350 does it (pure estimators: the entry of the
    last step of a block, weighted by that step's population; energies: block
    sums), starting from the lattice configuration of bench.py.  Default
    `jit_parallel=True`, the mode Proc.sampling always runs."""
    model, dmc = mrbp.model, mrbp.dmc
    spec = model.Spec(**kwargs)
    nop = spec.boson_number
    ssf_spec = dmc.SSFEstSpec(num_modes, as_pure_est=True,
                              pfw_num_time_steps=nts)
    den_spec = dmc.DensityEstSpec(num_bins, as_pure_est=True,
                                  pfw_num_time_steps=nts)
    sampling = dmc.Sampling(spec, dt, wmax, n_target,
                            num_walkers_control_factor=nwc, rng_seed=seed,
                            ssf_est_spec=ssf_spec, density_est_spec=den_spec,
                            jit_parallel=parallel)
    rng = np.random.default_rng(seed)
    confs = np.zeros((n_target, 2, nop))
    confs[:, 0, :] = (np.arange(nop)[None, :] + 0.25
                      + 0.15 * (rng.random((n_target, nop)) - 0.5))
    ini_state = sampling.build_state(confs)
    e_sum, w_sum, s_last, d_last, n_last = [], [], [], [], []
    for b, block in zip(range(burn + nblocks),
                        sampling.blocks(ini_state, nts, burn)):
        if b < burn:
            continue
        e_sum.append(block.iter_props.energy.sum())
        w_sum.append(block.iter_props.weight.sum())
        s_last.append(np.asarray(block.iter_ssf)[nts - 1].copy())
        d_last.append(np.asarray(block.iter_density)[nts - 1, :, 0].copy())
        n_last.append(float(block.iter_props.num_walkers[nts - 1]))
    e_sum, w_sum = np.array(e_sum), np.array(w_sum)
    s_last, d_last = np.array(s_last), np.array(d_last)
    n_last = np.array(n_last)
    e_mean = e_sum.sum() / w_sum.sum()
    _, e_err = _ref_reblock((e_sum - e_mean * w_sum) / w_sum.mean())
    out = dict(params=param_block(spec), ini_confs=confs, n_target=n_target,
               max_num_walkers=wmax, time_step=dt, nts=nts, nblocks=nblocks,
               burn=burn, nwc_factor=nwc, num_modes=num_modes,
               num_bins=num_bins, block_energy=e_sum, block_weight=w_sum,
               block_ssf_last=s_last, block_density_last=d_last,
               block_walkers_last=n_last, ref_energy_mean=e_mean,
               ref_energy_err=e_err, parallel=int(parallel))
    np.savez_compressed(os.path.join(GOLDEN, f'dmc_stat_pure_{name}.npz'),
                        **out)
    sk = s_last[:, :, 0].sum(axis=0) / n_last.sum() / nop
    print(f'dmc_stat_pure_{name}: E/N = {e_mean / nop:.6f} +- '
          f'{e_err / nop:.6f}; pure <|rho_k|^2>/N = {sk}; '
          f'density per bin = {d_last.sum(axis=0)[:4] / n_last.sum()}')


def gen_vmc_stat(mrbp, name, kwargs, *, move_spread, ns, nblocks, burn, seed,
                 num_modes=8):
    """A long single-chain reference VMC run (qmc_base/vmc.py:557-770):
    per-block means of E_L and of |rho_k|^2 with the error bars of the
    reference's own reblocking."""
    model, vmc = mrbp.model, mrbp.vmc
    spec = model.Spec(**kwargs)
    nop = spec.boson_number
    sampling = vmc.Sampling(spec, move_spread, rng_seed=seed,
                            ssf_est_spec=vmc.SSFEstSpec(num_modes))
    rng = np.random.default_rng(seed)
    conf = spec.get_sys_conf_buffer()
    conf[0, :] = rng.random(nop) * spec.supercell_size
    ini_state = sampling.build_state(conf)
    e_blk, s_blk, acc = [], [], []
    for b, block in zip(range(burn + nblocks),
                        sampling.blocks(ns, ini_state)):
        if b < burn:
            continue
        e_blk.append(block.iter_props.energy.mean())
        s_blk.append(np.asarray(block.iter_ssf)[:, :, 0].mean(axis=0))
        acc.append(block.accept_rate)
    e_blk, s_blk = np.array(e_blk), np.array(s_blk)
    e_mean, e_err = _ref_reblock(e_blk)
    sk = [_ref_reblock(s_blk[:, m]) for m in range(num_modes)]
    out = dict(params=param_block(spec), ini_conf=conf,
               move_spread=move_spread, ns=ns, nblocks=nblocks, burn=burn,
               num_modes=num_modes, block_energy=e_blk, block_ssf=s_blk,
               accept_rate=np.array(acc), ref_energy_mean=e_mean,
               ref_energy_err=e_err, ref_sk_mean=np.array([m for m, _ in sk]),
               ref_sk_err=np.array([e for _, e in sk]))
    np.savez_compressed(os.path.join(GOLDEN, f'vmc_stat_{name}.npz'), **out)
    print(f'vmc_stat_{name}: E/N = {e_mean / nop:.6f} +- {e_err / nop:.6f} '
          f'acc = {np.mean(acc):.4f}')


def gen_reblock(rng):
    """On-the-fly reblocking tables of the live reference
    (stats/reblock.py: on_the_fly_obj_create, on_the_fly_obj_data_update and
    the OTFObject statistics derived from them) for correlated series whose
    lengths are and are not powers of two."""
    import warnings
    from phd_qmclib.stats import reblock
    out = {}
    for tag, n, ncols in (('a', 100, 3), ('b', 64, 2), ('c', 12, 5),
                          ('d', 1, 1)):
        x = np.cumsum(rng.standard_normal((n, ncols)), axis=0) * 0.1 \
            + 10 * rng.random(ncols)
        otf = reblock.on_the_fly_obj_create(x)
        out[f'{tag}_series'] = x
        for f in otf.dtype.names:
            out[f'{tag}_{f}'] = otf[f]
    # accumulation over consecutive "blocks" and the statistics the
    # reference derives from an accumulated table
    blocks = [np.cumsum(rng.standard_normal(48)) * 0.05 + 3 for _ in range(5)]
    acc = reblock.on_the_fly_obj_create(blocks[0])
    for b in blocks[1:]:
        reblock.on_the_fly_obj_data_update(
            acc, reblock.on_the_fly_obj_create(b))
    out['acc_blocks'] = np.array(blocks)
    for f in acc.dtype.names:
        out[f'acc_{f}'] = acc[f]
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        obj = reblock.OTFObject(acc)
        out['acc_mean'] = obj.mean
        out['acc_errors'] = obj.errors
        out['acc_mean_eff_error'] = obj.mean_eff_error
    np.savez_compressed(os.path.join(GOLDEN, 'reblock_otf.npz'), **out)
    print('reblock_otf: mean', float(out['acc_mean']),
          'eff error', float(out['acc_mean_eff_error']))


def gen_cswf(mrbp, name, kwargs, rng, *, nconf, cutoffs):
    """Correlated-sampling objective of the wave-function optimiser
    (mrbp_qmc/model.py:817-942) on a fixed configuration set: ln|Psi| and
    E_L under each trial tbf_contact_cutoff from the reference's core
    functions, the variance from the reference's own weighed_variance."""
    import attr
    from phd_qmclib.qmc_base import jastrow
    model = mrbp.model
    spec = model.Spec(**kwargs)
    cf = model.core_funcs
    # one particle per randomly chosen cell, jittered inside the well: the
    # weights of such a set are comparable (iid uniform positions would
    # leave a single configuration with all the weight)
    nop, size = spec.boson_number, spec.supercell_size
    confs = np.zeros((nconf, 2, nop))
    for c in confs:
        cells = np.sort(rng.choice(int(size), nop, replace=False))
        c[0] = cells + spec.well_width * (0.5 + 0.5 * (rng.random(nop) - 0.5))
    ini = np.array([cf.wf_abs_log(c, *spec.cfc_spec) for c in confs])
    blocks, lns, ens, var = [], [], [], []
    for rm in cutoffs:
        trial = attr.evolve(spec, tbf_contact_cutoff=float(rm))
        cfc = trial.cfc_spec
        ln = np.array([cf.wf_abs_log(c, *cfc) for c in confs])
        en = np.array([cf.energy(c, *cfc) for c in confs])
        blocks.append(param_block(trial))
        lns.append(ln)
        ens.append(en)
        var.append(jastrow.CSWFOptimizer.weighed_variance(2 * (ln - ini), en))
    np.savez_compressed(
        os.path.join(GOLDEN, f'cswf_{name}.npz'),
        spec_keys=np.array(list(kwargs.keys())),
        spec_vals=np.array([float(v) for v in kwargs.values()]),
        params=param_block(spec), confs=confs, ini_lnpsi=ini,
        cutoffs=np.array(cutoffs, dtype=float), trial_params=np.array(blocks),
        lnpsi=np.array(lns), energy=np.array(ens), variance=np.array(var))
    print(f'cswf_{name}: variance = {np.array(var)}')


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    which = set(sys.argv[1:]) or {'model', 'step', 'branch', 'blocks', 'vmc',
                                  'stat', 'cswf', 'reblock'}
    mrbp = refshim.load()
    from phd_qmclib.mrbp_qmc import dmc, model, vmc  # noqa: F401
    draw_uniform, draw_dmc = numba_draws()
    if 'model' in which:
        for i, (name, kw) in enumerate(SPECS.items()):
            gen_model(mrbp, name, kw, np.random.default_rng(100 + i))
    if 'reblock' in which:
        gen_reblock(np.random.default_rng(700))
    if 'cswf' in which:
        gen_cswf(mrbp, 'll_n16', SPECS['ll_n16'], np.random.default_rng(600),
                 nconf=96, cutoffs=[0.05, 1.0, 2.5, 4.0, 6.0, 7.92])
        gen_cswf(mrbp, 'lat_n50', SPECS['lat_n50'],
                 np.random.default_rng(601), nconf=64,
                 cutoffs=[0.05, 3.0, 12.5, 20.0, 24.75])
        gen_cswf(mrbp, 'odd_n7', SPECS['odd_n7'], np.random.default_rng(602),
                 nconf=1500, cutoffs=[0.4, 3.3, 4.9])
    if 'step' in which:
        for i, name in enumerate(['ll_n16', 'lat_n50', 'defects_n20',
                                  'odd_n7']):
            gen_dmc_step(mrbp, name, SPECS[name],
                         np.random.default_rng(200 + i))
    if 'branch' in which:
        gen_branch(mrbp, np.random.default_rng(300), draw_uniform)
    if 'blocks' in which:
        gen_dmc_blocks(mrbp, 'll_n16', SPECS['ll_n16'],
                       np.random.default_rng(400), draw_dmc, n_ini=24,
                       wmax=40, dt=2e-3, nts=6, nblocks=2, num_modes=8,
                       num_bins=32, pure=True, seed=11)
        gen_dmc_blocks(mrbp, 'defects_n20_mixed', SPECS['defects_n20'],
                       np.random.default_rng(401), draw_dmc, n_ini=12,
                       wmax=24, dt=1e-3, nts=5, nblocks=2, num_modes=6,
                       num_bins=40, pure=False, seed=12)
    if 'vmc' in which:
        gen_vmc_blocks(mrbp, 'll_n16', SPECS['ll_n16'],
                       np.random.default_rng(500), draw_uniform,
                       move_spread=0.25, ns=40, nblocks=2, num_modes=8,
                       seed=5)
        gen_vmc_blocks(mrbp, 'defects_n20', SPECS['defects_n20'],
                       np.random.default_rng(501), draw_uniform,
                       move_spread=0.25 * (1 / 1.5), ns=40, nblocks=2,
                       num_modes=6, seed=6)
        gen_vmc_blocks(mrbp, 'ndf_lat_n50', SPECS['lat_n50'],
                       np.random.default_rng(502), draw_uniform,
                       move_spread=math.sqrt(2e-3), ns=24, nblocks=2,
                       num_modes=10, seed=7, gaussian=True)
    if 'statpure' in which:
        # BASELINE configs[2] model (N=50), 512 walkers, pure estimators
        # (tau = 20 after a burn-in of tau = 3: long against the relaxation
        # of the slowest S(k) mode, so that the blocked errors are honest)
        gen_dmc_stat_pure(mrbp, 'lat_n50', SPECS['lat_n50'], n_target=512,
                          wmax=640, dt=1e-3, nts=128, nblocks=160, burn=24,
                          seed=21)
    if 'statvmc50' in which:
        # BASELINE configs[1] particle number: one long reference chain
        gen_vmc_stat(mrbp, 'lat_n50', SPECS['lat_n50'], move_spread=0.25,
                     ns=4096, nblocks=96, burn=8, seed=3)
    if 'statpure100' in which:
        # BASELINE configs[3] model (N=100, dt of the headline bench)
        # (blocks of tau = 0.16: shorter ones leave the density bins of
        # neighbouring blocks correlated beyond what 96 blocks can resolve)
        gen_dmc_stat_pure(mrbp, 'lat_n100', SPECS['lat_n100'], n_target=256,
                          wmax=320, dt=6.25e-4, nts=256, nblocks=96, burn=16,
                          seed=22)
    if 'statpure200' in which:
        # BASELINE configs[4] model (N=200, deep lattice), S(k) and density
        gen_dmc_stat_pure(mrbp, 'deep_n200', SPECS['deep_n200'],
                          n_target=128, wmax=160, dt=1e-3, nts=128,
                          nblocks=80, burn=24, seed=23, num_modes=8,
                          num_bins=100)
    if 'stat' in which:
        gen_dmc_stat(mrbp, 'll_n16', SPECS['ll_n16'], n_target=512, wmax=640,
                     dt=2e-3, nts=256, nblocks=48, burn=12, seed=11)
        gen_dmc_stat(mrbp, 'lat_n16', dict(
            lattice_depth=5 * PI ** 2, lattice_ratio=1,
            interaction_strength=2, boson_number=16, supercell_size=16,
            tbf_contact_cutoff=4), n_target=512, wmax=640, dt=2e-3, nts=256,
            nblocks=48, burn=12, seed=11)
        gen_vmc_stat(mrbp, 'll_n16', SPECS['ll_n16'], move_spread=0.25,
                     ns=4096, nblocks=64, burn=4, seed=1)
        gen_vmc_stat(mrbp, 'defects_n20', SPECS['defects_n20'],
                     move_spread=0.25 * (1 / 1.5), ns=4096, nblocks=64,
                     burn=4, seed=2)


if __name__ == '__main__':
    main()
