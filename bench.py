#!/usr/bin/env python
"""Benchmark of the mrbp_qmc hot path on B200: walker-steps/s (DMC) or
chain-steps/s (VMC).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--config c4|c3_dmc50|c5_est|c2_vmc]

Workloads (BASELINE.json `configs`):
  c4        (default, the headline) configs[3]: DMC N = L = 100, V0 = 5 pi^2,
            g = 2, r_m = L/4, dt = 6.25e-4, 1.25e5 target walkers PER GPU (1e6
            on 8 GPUs; capacity 1.25x), kappa = 0.5, blocks of 128 time steps
  c3_dmc50  configs[2]: DMC N = 50, 1e4 target walkers, dt = 1e-3, blocks of 512
  c5_est    configs[4]: DMC N = 200 deep lattice (V0 = 20 pi^2), 3.125e4 target
            walkers per GPU (2.5e5 on 8), dt = 1e-3, pure S(k) with M = 400 and
            pure density with B = 6400 evaluated every step, blocks of 64
  c2_vmc    configs[1]: VMC N = 50, 1e5 independent chains per GPU, energy and
            S(k) (M = 50) estimators, blocks of 256 steps
One bench "step" = one block = one `next()` of the reference's
`Sampling.blocks()` = one qmcb_dmc_run_block / qmcb_vmc_run_block call.  Weak
scaling: the per-GPU population is fixed as N grows.

One JSON line on stdout (rank 0):
  value  units/s of the whole job, population resident in HBM
  e2e    same metric through the C ABI with HOST buffers: every step copies
         the whole population host->device (pinned), runs the block, and
         copies the evolved population and the block's results back
  roofline      the dominant kernel against the fp64 DFMA peak measured in
                this run (MEASURED_PEAKS.json has no fp64 entry) and nominal
  cpu_baseline  the reference's own Numba implementation on the host cores
                (kind "reference"; the C port of oracle/ with an explicit
                reason only when the reference cannot run)

`--impl reference` times the reference (`mrbp_qmc.{dmc,vmc}.Sampling.blocks()`
from baseline/_ref, through oracle/refshim.py) alone on the host cores.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NWC = 0.5
CAP_FACTOR = 1.25
SEED = 7
PI2 = math.pi ** 2

CONFIGS = {
    'c4': dict(
        kind='dmc', nop=100, v0=5 * PI2, g=2.0, dt=6.25e-4, walkers=125000,
        nts=128, modes=0, bins=0, metric='dmc_walker_steps_per_sec',
        unit='walker-steps/s', baseline_config='configs[3] shard',
        ref=dict(nw=4096, nts=16)),
    'c3_dmc50': dict(
        kind='dmc', nop=50, v0=5 * PI2, g=2.0, dt=1e-3, walkers=10000,
        nts=512, modes=0, bins=0, metric='dmc_walker_steps_per_sec',
        unit='walker-steps/s', baseline_config='configs[2]',
        ref=dict(nw=4096, nts=16)),
    'c5_est': dict(
        kind='dmc', nop=200, v0=20 * PI2, g=2.0, dt=1e-3, walkers=31250,
        nts=64, modes=400, bins=6400, metric='dmc_walker_steps_per_sec',
        unit='walker-steps/s', baseline_config='configs[4] shard',
        ref=dict(nw=512, nts=4)),
    'c2_vmc': dict(
        kind='vmc', nop=50, v0=5 * PI2, g=4.0, chains=100000, ns=256,
        modes=50, metric='vmc_chain_steps_per_sec', unit='chain-steps/s',
        baseline_config='configs[1]', ref=dict(ns=4096)),
}


def flops_per_walker_step(n):
    """Algorithmic work model of SURVEY.md 8(d) (DMC step)."""
    return 58 * n * (n - 1) / 2 + 113 * n + 35


def ssf_flops(n, m):
    """S(k) per evaluated configuration."""
    return n * (40 + 8 * m) + 3 * m


def vmc_flops_per_chain_step(n, m, accept):
    """Same convention for a VMC step: every step moves N
    particles (8 flop each) and evaluates ln|Psi| (per unordered pair: 6 for
    the distance and argument + 40 for the trigonometric function + 40 for
    its logarithm; 90 per particle for ln f1); an accepted step adds E_L
    (58 per pair + 50 per particle) and rho_k."""
    pairs = n * (n - 1) / 2
    return (8 * n + 86 * pairs + 90 * n
            + accept * (58 * pairs + 50 * n + (ssf_flops(n, m) if m else 0)))


def bytes_per_walker_step(n):
    return 32 * n + 32


def spec_kwargs(cfg):
    n = cfg['nop']
    return dict(lattice_depth=cfg['v0'], lattice_ratio=1.0,
                interaction_strength=cfg['g'], boson_number=n,
                supercell_size=float(n), tbf_contact_cutoff=0.25 * n)


def model_spec(cfg):
    from phd_qmclib_b200 import model
    return model.Spec(**spec_kwargs(cfg))


def initial_confs(nw, nop, seed):
    """Walkers near the Mott-like ground state: one boson per lattice well
    (well = [0, 1/2) of each unit cell) with a small random offset, so the
    population equilibrates within the warm-up blocks."""
    rng = np.random.default_rng(seed)
    ini = np.zeros((nw, 2, nop))
    ini[:, 0] = (np.arange(nop)[None, :] + 0.25
                 + 0.15 * (rng.random((nw, nop)) - 0.5))
    return ini


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,'
         'clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--id={self.index}',
                 f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                 '-lms', '100'], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [],
                    'note': 'nvidia-smi unavailable'}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap']
        for ts, line in self.rows:
            if t0 is not None and not (t0 <= ts <= t1 + 0.2):
                continue
            f = [x.strip() for x in line.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
                pw.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nme)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [],
                    'note': 'no samples'}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)),
                'power_w_max': float(max(pw)), 'samples': len(sm),
                'reasons': sorted(reasons)}


def workload_config(name, cfg, args, world):
    n = cfg['nop']
    if cfg['kind'] == 'vmc':
        return {
            'workload': f'mrbp_qmc VMC N={n} (BASELINE {cfg["baseline_config"]}'
                        f'): V0={cfg["v0"] / PI2:g}pi^2 g={cfg["g"]:g} L={n} '
                        f'r_m={n // 4} move_spread=0.25 well widths, energy + '
                        f'S(k) M={cfg["modes"]} every step',
            'bench_config': name, 'boson_number': n,
            'chains_per_gpu': args.walkers,
            'global_chains': args.walkers * world,
            'steps_per_step': args.nts,
            'parallelism': f'chains sharded over {world} GPU(s), no '
                           f'collective',
            'l2_policy': 'chain state lives in registers for the whole '
                         'block; per-chain sums (%.0f MB) written once per '
                         'block' % (args.walkers * (2 + 3 * cfg['modes']) * 8
                                    / 1e6),
        }
    est = ('off in the timed region' if not (cfg['modes'] or cfg['bins']) else
           f'pure S(k) M={cfg["modes"]} + pure density B={cfg["bins"]} '
           f'every step, inside the timed region')
    return {
        'workload': f'mrbp_qmc DMC N={n} (BASELINE {cfg["baseline_config"]}): '
                    f'V0={cfg["v0"] / PI2:g}pi^2 g={cfg["g"]:g} L={n} '
                    f'r_m={n // 4} dt={cfg["dt"]:g} kappa={NWC}',
        'bench_config': name, 'boson_number': n,
        'target_walkers_per_gpu': args.walkers,
        'capacity_per_gpu': int(args.walkers * CAP_FACTOR),
        'global_target_walkers': args.walkers * world,
        'time_steps_per_step': args.nts,
        'estimators': est,
        'parallelism': f'walkers sharded over {world} GPU(s)',
        'l2_policy': _l2_policy(cfg, args),
    }


def _l2_policy(cfg, args):
    n, M = cfg['nop'], cfg['modes']
    cap = args.walkers * CAP_FACTOR
    txt = 'working set (2 x %.0f MB walker buffers' % (cap * 16 * n / 1e6)
    if M:
        txt += ' + %.0f MB S(k) rows' % (2 * cap * 24 * M / 1e6)
    if cap * 32 * n + 2 * cap * 24 * M > 126e6:
        return txt + ') exceeds the 126 MB L2'
    return txt + (') fits the 126 MB L2 (the configuration BASELINE names is '
                  'this small): L2 flushed before every timed block by a '
                  '256 MB memset outside the timed events')


# ---------------------------------------------------------------------------
# CPU arm: the reference's Numba implementation on the host cores; the C port
# of oracle/ only as a fallback with an explicit reason
# ---------------------------------------------------------------------------
def _oracle_path():
    p = os.path.join(ROOT, 'oracle')
    if p not in sys.path:
        sys.path.insert(0, p)


def cpu_port_units_per_s(cfg, budget_s=12.0, threads=None):
    """Oracle port (OpenMP over walkers, like the reference's prange) on a
    bounded sample of the workload.  Returns a cpu_baseline-shaped dict."""
    _oracle_path()
    import oracle
    from phd_qmclib_b200 import model
    oracle.lib()
    cores = threads or os.cpu_count() or 1
    oracle.set_num_threads(cores)
    p = model.param_block(model_spec(cfg))
    n = cfg['nop']
    if cfg['kind'] == 'vmc':
        nch, ns = 4 * cores, 64
        cur = initial_confs(nch, n, 0)
        ln = oracle.model_eval(p, cur, want=('lnpsi',))['lnpsi']
        spread = 0.25 * model_spec(cfg).well_width
        eprev, sprev = np.zeros(nch), np.zeros((nch, cfg['modes'], 3))
        oracle.vmc_block(p, 1, spread, 0.0, float(n), cur, ln, eprev, sprev,
                         cfg['modes'], 2, 0, True)
        t0 = time.perf_counter()
        oracle.vmc_block(p, 1, spread, 0.0, float(n), cur, ln, eprev, sprev,
                         cfg['modes'], ns, 1, False)
        dt = time.perf_counter() - t0
        return dict(value=nch * ns / dt, unit=cfg['unit'], cores=cores,
                    kind='port', seconds=dt,
                    sample=f'{nch} chains x {ns} steps, oracle/qmc_oracle.c '
                           f'with OpenMP')
    nw = cfg['ref']['nw']
    cap = int(nw * CAP_FACTOR)
    st = oracle.DMCState(p, initial_confs(nw, n, 11), cap)
    t0 = time.perf_counter()
    it = st.run_block(SEED, cfg['dt'], nw, NWC, 1, 0.0, float(n))
    t1 = time.perf_counter() - t0
    nts = int(max(2, min(256, budget_s / max(t1, 1e-4))))
    t0 = time.perf_counter()
    it = st.run_block(SEED, cfg['dt'], nw, NWC, nts, 0.0, float(n))
    dt = time.perf_counter() - t0
    ws = float(it['num_walkers'].sum())
    return dict(value=ws / dt, unit=cfg['unit'], cores=cores, kind='port',
                seconds=dt,
                sample=f'{nw} target / {cap} capacity walkers x {nts} time '
                       f'steps of the bench workload (estimators off), '
                       f'oracle/qmc_oracle.c with OpenMP')


def reference_runner(cfg, parallel=True, shrink=1):
    """The LIVE reference under oracle/refshim.py (protocol of SURVEY.md 8d /
    BASELINE.md section 3) as an object whose `step()` times one more block.
    Raises RuntimeError with the reason when it cannot run here."""
    _oracle_path()
    import ref_arm
    ok, why = ref_arm.probe()
    if not ok:
        raise RuntimeError(why)
    kw = spec_kwargs(cfg)
    if cfg['kind'] == 'vmc':
        from phd_qmclib_b200 import model
        return ref_arm.VmcRun(
            kw, move_spread=0.25 * model.Spec(**kw).well_width,
            ns=cfg['ref']['ns'], num_modes=cfg['modes'])
    nw = max(64, cfg['ref']['nw'] // shrink)
    return ref_arm.DmcRun(kw, nw=nw, cap=int(nw * CAP_FACTOR), dt=cfg['dt'],
                          nwc=NWC, nts=cfg['ref']['nts'], parallel=parallel,
                          num_modes=cfg['modes'], num_bins=cfg['bins'])


class PortRunner:
    """Fallback: the C port of oracle/ (kind "port"), with the reason."""

    def __init__(self, cfg, reason):
        self.cfg, self.reason = cfg, reason
        self.sample, self.cores = '', os.cpu_count() or 1

    def step(self):
        r = cpu_port_units_per_s(self.cfg, budget_s=4.0)
        self.sample, self.cores = r['sample'], r['cores']
        return r['value'] * r['seconds'], r['seconds']


def run_reference(name, cfg, args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    cb = {'unit': cfg['unit']}
    try:
        run = reference_runner(cfg)
        _oracle_path()
        import ref_arm
        cb.update(kind='reference', threading_layer=run.layer,
                  numba=run.numba, cpu_model=ref_arm.cpu_model(),
                  host_cores=os.cpu_count(),
                  jit_and_first_block_s=run.first_block_s)
    except Exception as exc:       # noqa: BLE001 - any failure is a reason
        run = PortRunner(cfg, 'reference (Numba) arm unavailable: '
                         f'{exc.__class__.__name__}: {exc}'[:400])
        cb.update(kind='port', reason=run.reason)
    # each bench "step" is one block of the bounded sample; the JIT
    # compilation of the reference (~1-2 min) happened in the constructor
    per_step = []
    for i in range(max(1, args.steps + args.warmup)):
        units, secs = run.step()
        if i >= args.warmup or args.steps + args.warmup == 0:
            per_step.append((units, secs))
    value = float(sum(u for u, _ in per_step) / sum(t for _, t in per_step))
    ms = float(np.mean([t for _, t in per_step])) * 1e3
    cb.update(value=value, cores=run.cores, sample=run.sample)
    if cb['kind'] == 'reference' and cfg['kind'] == 'dmc':
        # one core, for the record (its own JIT compilation; smaller sample)
        try:
            ser = reference_runner(cfg, parallel=False, shrink=8)
            u, t = ser.step()
            cb['serial'] = {'value': u / t, 'cores': 1, 'sample': ser.sample}
        except Exception as exc:    # noqa: BLE001
            cb['serial'] = {'value': None, 'reason': str(exc)[:200]}
    line = {
        'impl': 'reference', 'metric': cfg['metric'],
        'value': value, 'unit': cfg['unit'], 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(name, cfg, args, 1),
        'cpu_baseline': cb,
        'e2e': {'value': value, 'unit': cfg['unit'],
                'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def cpu_baseline_subprocess(name):
    """The cpu_baseline leg of the GPU arm runs the reference arm in a child
    process (its own OpenMP/Numba thread pools, none of torch's) and takes
    the cpu_baseline object of its line."""
    cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference',
           '--config', name, '--steps', '2', '--warmup', '0']
    env = {k: v for k, v in os.environ.items()
           if k not in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE')}
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, env=env,
                             timeout=900)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith('{'):
                return json.loads(ln)['cpu_baseline']
        return {'value': None, 'kind': 'unavailable',
                'reason': (out.stderr or 'no output')[-300:]}
    except Exception as exc:       # noqa: BLE001
        return {'value': None, 'kind': 'unavailable', 'reason': str(exc)}


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def bind_to_gpu_numa_node(index):
    """One process per GPU: run on the host cores next to that GPU (NVML's
    CPU affinity of the device), so that the pinned staging buffers of the
    end-to-end leg are first touched on the GPU's own NUMA node instead of
    all eight processes sharing one node's memory and PCIe root."""
    try:
        import pynvml
        pynvml.nvmlInit()
        hdl = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(hdl, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64)
                if (m >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f'{len(cpus)} cores next to GPU {index}'
    except Exception as exc:        # noqa: BLE001 - best effort
        return f'not bound ({exc.__class__.__name__})'
    return 'not bound'


class Dist:
    """torch.distributed plumbing of the bench (barrier, max/sum over ranks)."""

    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        if not torch.cuda.is_available():
            raise SystemExit('bench.py: no CUDA device (there is no CPU '
                             'fallback for the product arm; use --impl '
                             'reference)')
        torch.cuda.set_device(self.local)
        self.affinity = bind_to_gpu_numa_node(self.local) \
            if self.world > 1 else None
        self.dist = None
        self.stdout_fd = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            # stdout carries exactly one JSON line.  NCCL writes its version
            # banner (NCCL_DEBUG=VERSION, the setting of the GPU boxes) to
            # file descriptor 1 from C: the environment is left alone, fd 1
            # points at stderr for the duration of the run, and the line
            # goes to the saved descriptor.
            sys.stdout.flush()
            self.stdout_fd = os.dup(1)
            os.dup2(2, 1)
            dist.init_process_group(
                'nccl', device_id=torch.device('cuda', self.local))
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def _red(self, x, op):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device='cuda')
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._red(x, self.dist.ReduceOp.MAX) if self.dist else x

    def sum(self, x):
        return self._red(x, self.dist.ReduceOp.SUM) if self.dist else x

    def gather(self, x):
        if self.dist is None:
            return [x]
        t = self.torch.zeros(self.world, dtype=self.torch.float64,
                             device='cuda')
        t[self.rank] = x
        self.dist.all_reduce(t)
        return [float(v) for v in t.tolist()]

    def emit(self, line):
        data = (json.dumps(line) + '\n').encode()
        if self.stdout_fd is None:
            sys.stdout.write(data.decode())
            sys.stdout.flush()
        else:
            os.write(self.stdout_fd, data)

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        return {}


NOMINAL_FP64_TF = 148 * 64 * 2 * 1.965e9 / 1e12    # 37.2


def roofline_fp64(kernel, ach_tf, peak_sust, peak_burst, extra):
    r = {
        'kernel': kernel, 'bound': 'fp64', 'achieved': ach_tf,
        'peak': peak_sust, 'unit': 'TFLOP/s', 'frac': ach_tf / peak_sust,
        'peak_source': 'DFMA microbenchmark measured in this run, sustained '
                       '1.5 s (MEASURED_PEAKS.json has no fp64 figure)',
        'peak_burst': peak_burst, 'frac_of_burst': ach_tf / peak_burst,
        'peak_nominal': NOMINAL_FP64_TF,
        'frac_of_nominal': ach_tf / NOMINAL_FP64_TF,
    }
    r.update(extra)
    return r


def committed_ncu_view(name):
    """Executed-instruction view of the dominant kernel from the committed
    ncu capture of this config (not measured in this run)."""
    path = os.path.join(ROOT, 'profiles', 'kernel_ncu_view.json')
    try:
        return json.load(open(path)).get(name)
    except Exception:
        return None


def run_dmc(name, cfg, args, D):
    import torch
    from phd_qmclib_b200 import engine, _lib
    world, rank, local = D.world, D.rank, D.local
    n = cfg['nop']
    nw, nts = args.walkers, args.nts
    cap = int(nw * CAP_FACTOR)
    M, NB = cfg['modes'], cfg['bins']
    est = bool(M or NB)
    spec = model_spec(cfg)
    eng = engine.Engine(spec, device=local)
    dp = eng.dmc_params(cfg['dt'], cap * world, nw * world, NWC, SEED, 0.0,
                        float(n), local_capacity=cap,
                        ssf=(M, True, nts) if M else None,
                        density=(NB, True, nts) if NB else None)
    if world > 1:
        eng.comm_init_torch(D.dist, rank, world)
    eng.dmc_init(dp, initial_confs(nw, n, 100 + rank),
                 global_slot_offset=rank * cap)
    stream = torch.cuda.ExternalStream(eng.stream)

    # fp64 roofline denominator, measured on this GPU, now
    peak_burst = engine.measure_fp64_peak(local)
    peak_sust = engine.measure_fp64_peak(local, sustained_seconds=1.5)

    # small working sets (c3_dmc50: 2 x 10 MB) would stay L2-resident between
    # blocks: flush with a buffer larger than L2 between timed blocks
    ws_bytes = cap * 32 * n + 2 * cap * 24 * M
    flush = None
    if ws_bytes <= 126e6:
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8,
                            device='cuda')

    den = np.zeros((nts, NB)) if NB else None
    ssf = np.zeros((nts, M, 3)) if M else None
    series = dict(energy=np.zeros(nts), weight=np.zeros(nts),
                  num_walkers=np.zeros(nts, dtype=np.uint64),
                  ref_energy=np.zeros(nts), accum_energy=np.zeros(nts))

    def block():
        eng.dmc_run_block(nts, eval_estimators=est, out=series, density=den,
                          ssf=ssf)

    # ---- device-resident timing ------------------------------------------
    # Per-launch events (set_profiling) give the step kernel's time inside
    # the timed region, but rule out the CUDA graph a launch-bound block
    # replays; small populations are therefore timed as the product runs
    # them and the kernel time comes from extra profiled blocks afterwards.
    profile_in_region = world > 1 or est or cap * n >= 2000000
    eng.set_profiling(profile_in_region)
    moved = 0
    for _ in range(args.warmup):
        block()
        if world > 1:
            eng.dmc_rebalance()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    D.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    ws_total, kern_ms, launches, dev_ms, kern_ws = 0.0, 0.0, 0, 0.0, 0.0
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    if flush is None:
        e0.record(stream)
    for _ in range(args.steps):
        if flush is not None:
            with torch.cuda.stream(stream):
                flush.zero_()
            e0.record(stream)
        block()
        st = eng.last_block_stats()
        kern_ms += st['step_kernel_ms']
        launches += st['launches']
        kern_ws += float(series['num_walkers'].sum())
        ws_total += float(series['num_walkers'].sum())   # GLOBAL if world > 1
        if world > 1:
            # order-preserving neighbour shifts, once per block, timed
            moved += eng.dmc_rebalance()
        if flush is not None:
            e1.record(stream)
            torch.cuda.synchronize()
            dev_ms += e0.elapsed_time(e1)
    if flush is None:
        e1.record(stream)
    D.barrier(); torch.cuda.synchronize()
    t1 = time.perf_counter()
    if flush is None:
        dev_ms = e0.elapsed_time(e1)
    local_ms = dev_ms
    dev_ms = D.max(dev_ms)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    value = ws_total / (dev_ms * 1e-3)
    ms_per_step = dev_ms / args.steps
    e_per_particle = float(series['energy'][-1] / series['weight'][-1] / n)
    n_local = float(eng.dmc_scalars().num_walkers)
    per_rank = {'walkers': D.gather(n_local),
                'step_kernel_ms_per_launch': D.gather(
                    kern_ms / (args.steps * nts) if profile_in_region
                    else float('nan')),
                'block_ms': D.gather(local_ms / args.steps)}

    kern_blocks = args.steps
    if not profile_in_region:
        eng.set_profiling(True)
        kern_ms, kern_ws, kern_blocks = 0.0, 0.0, 3
        for _ in range(kern_blocks):
            block()
            kern_ms += eng.last_block_stats()['step_kernel_ms']
            kern_ws += float(series['num_walkers'].sum())

    # roofline of the step kernel (this rank's launches, this rank's walkers)
    ws_rank = kern_ws / world
    F = flops_per_walker_step(n)
    ach_tf = ws_rank * F / (kern_ms * 1e-3) / 1e12
    hbm_gbs = ws_rank * bytes_per_walker_step(n) / (kern_ms * 1e-3) / 1e9
    view = committed_ncu_view(name) or {}
    traffic = None
    if view.get('dram_bytes_per_walker') is not None:
        traffic = view['dram_bytes_per_walker'] * ws_rank / (kern_blocks * nts)
    peaks = measured_peaks()
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    roofline = roofline_fp64('dmc_step_kernel', ach_tf, peak_sust, peak_burst, {
        'flop_per_walker_step': F,
        'avg_launch_ms': kern_ms / (kern_blocks * nts),
        'step_kernel_share_of_step': (kern_ms / kern_blocks) / ms_per_step,
        'kernel_timing': ('CUDA events around every launch inside the timed '
                          'region' if profile_in_region else
                          f'CUDA events around every launch of {kern_blocks} '
                          f'extra blocks after the timed region (the timed '
                          f'blocks replay a CUDA graph)'),
        'traffic': traffic, 'ncu': view or None,
        'hbm': {'achieved': hbm_gbs, 'peak': hbm_peak, 'unit': 'GB/s',
                'frac': hbm_gbs / hbm_peak,
                'bytes_per_walker_step': bytes_per_walker_step(n)},
    })

    # ---- end to end through the C ABI with host buffers -------------------
    eng.set_profiling(False)
    nx = eng.dmc_get_next()
    n_live = int(nx['scalars'].num_walkers)
    h_confs = engine.pinned_empty((cap, 2, n))
    h_energy = engine.pinned_empty((cap,))
    h_weight = engine.pinned_empty((cap,))
    h_slot = engine.pinned_empty((cap,))
    h_confs[:n_live] = nx['confs']; h_energy[:n_live] = nx['energy']
    h_weight[:n_live] = nx['weight']; h_slot[:] = nx['slot_energy']
    sc = nx['scalars']
    h2d = d2h = 0
    e2e_ws = 0.0
    est_bytes = nts * (3 * M + NB) * 8

    def e2e_step(sc, n_live):
        eng.dmc_set_state(dp, h_confs[:n_live], h_energy[:n_live],
                          h_weight[:n_live], sc, slot_energy=h_slot,
                          global_slot_offset=rank * cap)
        block()
        if world > 1:
            eng.dmc_rebalance()
        sc2 = eng.dmc_get_next_into(h_confs, h_energy, h_weight, h_slot)
        return sc2, int(sc2.num_walkers)

    sc, n_live = e2e_step(sc, n_live)        # warm-up
    D.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        n_in = n_live
        sc, n_live = e2e_step(sc, n_live)
        h2d += n_in * (2 * n + 2) * 8 + cap * 8
        d2h += n_live * (2 * n + 2) * 8 + cap * 8 + nts * 5 * 8 + est_bytes
        e2e_ws += float(series['num_walkers'].sum())
    D.barrier(); torch.cuda.synchronize()
    e2e_wall = D.max(time.perf_counter() - t0)
    e2e = {'value': e2e_ws / e2e_wall, 'unit': cfg['unit'],
           'h2d_bytes_per_step': int(D.sum(h2d) / args.steps),
           'd2h_bytes_per_step': int(D.sum(d2h) / args.steps),
           'timing': 'host wall clock around K x (set_state from pinned '
                     'host + run_block + get_next to pinned host), max over '
                     'ranks',
           'ms_per_step': e2e_wall * 1e3 / args.steps}
    hits = int(eng.dmc_scalars().capacity_hits)
    eng.close()
    if rank != 0:
        return None
    return {
        'metric': cfg['metric'], 'value': value, 'unit': cfg['unit'],
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': workload_config(name, cfg, args, world),
        'per_gpu_value': value / world,
        'energy_per_particle_last_step': e_per_particle,
        'roofline': roofline, 'e2e': e2e, 'gpu_launches': int(launches),
        'clocks': clocks, 'per_rank': per_rank,
        'rebalanced_walkers_rank0': int(moved), 'capacity_hits_rank0': hits,
        'host_affinity_rank0': D.affinity,
        'l2_flush': flush is not None,
        'lib': os.path.relpath(_lib.LIB_PATH, ROOT),
    }


def run_vmc(name, cfg, args, D):
    import torch
    from phd_qmclib_b200 import engine, _lib
    world, rank, local = D.world, D.rank, D.local
    n, M = cfg['nop'], cfg['modes']
    nch, ns = args.walkers, args.nts
    spec = model_spec(cfg)
    spread = 0.25 * spec.well_width
    eng = engine.Engine(spec, device=local)
    ini = initial_confs(nch, n, 200 + rank)
    eng.vmc_init(ini, spread, 1, 0.0, float(n), ssf_num_modes=M,
                 chain_offset=rank * nch)
    stream = torch.cuda.ExternalStream(eng.stream)
    peak_burst = engine.measure_fp64_peak(local)
    peak_sust = engine.measure_fp64_peak(local, sustained_seconds=1.5)

    for _ in range(args.warmup):
        eng.vmc_run_block(ns, series=False, sums=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    D.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record(stream)
    kern_ms, acc = 0.0, 0.0
    for _ in range(args.steps):
        # the per-chain sums stay on the device (accumulated by the kernel)
        o = eng.vmc_run_block(ns, series=False, sums=False)
        kern_ms += eng.last_block_stats()['total_ms']
        acc += float(o['accept_rate'].mean())
    e1.record(stream)
    D.barrier(); torch.cuda.synchronize()
    t1 = time.perf_counter()
    dev_ms = D.max(e0.elapsed_time(e1))
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    units = float(nch) * ns * args.steps * world
    value = units / (dev_ms * 1e-3)
    acc /= args.steps
    F = vmc_flops_per_chain_step(n, M, acc)
    ach_tf = (units / world) * F / (kern_ms * 1e-3) / 1e12
    roofline = roofline_fp64('vmc_block_kernel', ach_tf, peak_sust,
                             peak_burst, {
        'flop_per_chain_step': F, 'accept_rate': acc,
        'avg_launch_ms': kern_ms / args.steps,
        'kernel_share_of_step': kern_ms / dev_ms,
        'traffic': None, 'ncu': committed_ncu_view(name),
    })

    o = eng.vmc_run_block(ns, series=False, sums=True)
    esum = float(o['sum_energy'][:, 0].mean() / ns / n)

    # ---- end to end: chains from host, block, per-chain sums to host ------
    confs, _ = eng.vmc_get_state()
    h_confs = engine.pinned_empty((nch, 2, n))
    h_confs[:] = confs

    # results land in page-locked arrays allocated once (the chains come
    # back into the buffer the next block is initialised from)
    h_out = {'accept_rate': engine.pinned_empty((nch,)),
             'sum_energy': engine.pinned_empty((nch, 2))}
    if M:
        h_out['sum_ssf'] = engine.pinned_empty((nch, M, 3))
    h_ln = engine.pinned_empty((nch,))

    def e2e_step():
        eng.vmc_init(h_confs, spread, 1, 0.0, float(n), ssf_num_modes=M,
                     chain_offset=rank * nch)
        o = eng.vmc_run_block(ns, series=False, sums=True, out=h_out)
        eng.vmc_get_state(out=(h_confs, h_ln))
        return o

    e2e_step()
    D.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    D.barrier(); torch.cuda.synchronize()
    wall = D.max(time.perf_counter() - t0)
    h2d = nch * 2 * n * 8
    d2h = nch * (2 * n + 1) * 8 + nch * (1 + 2 + 3 * M) * 8
    e2e = {'value': units / wall, 'unit': cfg['unit'],
           'h2d_bytes_per_step': int(h2d * world),
           'd2h_bytes_per_step': int(d2h * world),
           'timing': 'host wall clock around K x (vmc_init from pinned host '
                     '+ run_block with per-chain sums to pinned host + '
                     'get_state to pinned host), '
                     'max over ranks; the first step of a re-initialised '
                     'block re-evaluates the initial configuration',
           'ms_per_step': wall * 1e3 / args.steps}
    eng.close()
    if rank != 0:
        return None
    return {
        'metric': cfg['metric'], 'value': value, 'unit': cfg['unit'],
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dev_ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': workload_config(name, cfg, args, world),
        'per_gpu_value': value / world,
        'energy_per_particle_last_block': esum, 'accept_rate': acc,
        'roofline': roofline, 'e2e': e2e, 'gpu_launches': int(args.steps),
        'clocks': clocks, 'lib': os.path.relpath(_lib.LIB_PATH, ROOT),
    }


def run_b200(name, cfg, args):
    D = Dist()
    line = (run_vmc if cfg['kind'] == 'vmc' else run_dmc)(name, cfg, args, D)
    if D.rank == 0:
        line['cpu_baseline'] = None
        if D.world == 1 and not args.no_cpu:
            line['cpu_baseline'] = cpu_baseline_subprocess(name)
        D.emit(line)
    D.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='c4', choices=sorted(CONFIGS))
    ap.add_argument('--walkers', type=int, default=None,
                    help='target walkers (chains) per GPU')
    ap.add_argument('--nts', type=int, default=None,
                    help='time steps per bench step (block)')
    ap.add_argument('--no-cpu', action='store_true',
                    help='skip the cpu_baseline leg')
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.walkers is None:
        args.walkers = cfg.get('walkers', cfg.get('chains'))
    if args.nts is None:
        args.nts = cfg.get('nts', cfg.get('ns'))
    if args.impl == 'reference':
        return run_reference(args.config, cfg, args)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.gpus != world and args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
               f'--nproc-per-node={args.gpus}', '--master-addr', '127.0.0.1',
               '--master-port', '29511', os.path.abspath(__file__)] \
            + sys.argv[1:]
        return subprocess.call(cmd)
    return run_b200(args.config, cfg, args)


if __name__ == '__main__':
    sys.exit(main())
