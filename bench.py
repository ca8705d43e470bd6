#!/usr/bin/env python
"""Headline benchmark: DMC walker-steps/s of the mrbp_qmc hot path (N=100).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[3], SURVEY.md 8d "C4"): multi-rods Bose gas,
N = L = 100, V0 = 5 pi^2, g = 2, r_m = L/4, dt = 6.25e-4, 1.25e5 target
walkers PER GPU (1e6 on 8 GPUs; capacity 1.25x), kappa = 0.5.  One bench
"step" = one block of `nts` DMC time steps = one `next()` of the reference's
`Sampling.blocks()` = one qmcb_dmc_run_block call.  Weak scaling: per-GPU
population is fixed as N grows.

Lines printed (rank 0, one JSON object):
  value  walker-steps/s of the whole job, walkers resident in HBM
  e2e    same metric through the C ABI with HOST buffers: every step copies
         the whole population host->device (pinned), runs the block, and
         copies the evolved population and the per-step series back
  roofline      the fused step kernel against the fp64 DFMA peak measured in
                this run (MEASURED_PEAKS.json has no fp64 entry)
  cpu_baseline  the oracle port of the reference algorithm on the host cores

`--impl reference` times that oracle port alone (the reference is Python +
Numba and does not exist on the GPU box).
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

NOP = 100
TIME_STEP = 6.25e-4
NWC = 0.5
CAP_FACTOR = 1.25
SEED = 7


def flops_per_walker_step(n):
    """Algorithmic work model of SURVEY.md 8(d)."""
    return 58 * n * (n - 1) / 2 + 113 * n + 35


def bytes_per_walker_step(n):
    return 32 * n + 32


def model_spec():
    from phd_qmclib_b200 import model
    return model.Spec(5 * math.pi ** 2, 1, 2, NOP, NOP, 0.25 * NOP)


def initial_confs(nw, seed):
    """Walkers near the Mott-like ground state: one boson per lattice well
    (well = [0, 1/2) of each unit cell) with a small random offset, so the
    population equilibrates within the warm-up blocks."""
    rng = np.random.default_rng(seed)
    ini = np.zeros((nw, 2, NOP))
    ini[:, 0] = (np.arange(NOP)[None, :] + 0.25
                 + 0.15 * (rng.random((nw, NOP)) - 0.5))
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


# ---------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm on the host cores
# ---------------------------------------------------------------------------
def cpu_port_walker_steps_per_s(budget_s=12.0, nw=4096, threads=None):
    """Oracle DMC (OpenMP over walkers, like the reference's prange) on a
    bounded sample of the bench workload.  Returns (value, cores, sample)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import oracle
    from phd_qmclib_b200 import model
    oracle.lib()
    cores = threads or os.cpu_count() or 1
    oracle.set_num_threads(cores)
    p = model.param_block(model_spec())
    cap = int(nw * CAP_FACTOR)
    st = oracle.DMCState(p, initial_confs(nw, 11), cap)
    t0 = time.perf_counter()
    it = st.run_block(SEED, TIME_STEP, nw, NWC, 1, 0.0, float(NOP))
    t1 = time.perf_counter() - t0
    nts = int(max(2, min(256, budget_s / max(t1, 1e-4))))
    t0 = time.perf_counter()
    it = st.run_block(SEED, TIME_STEP, nw, NWC, nts, 0.0, float(NOP))
    dt = time.perf_counter() - t0
    ws = float(it['num_walkers'].sum())
    sample = (f'{nw} target / {cap} capacity walkers x {nts} time steps of '
              f'the bench workload, oracle/qmc_oracle.c with OpenMP')
    return ws / dt, cores, sample, dt


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    per_step = []
    sample = ''
    # each "step" is a bounded sample sized to keep the whole run short
    budget = max(1.5, min(10.0, 100.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        v, cores, sample, dt = cpu_port_walker_steps_per_s(budget_s=budget)
        if i >= args.warmup:
            per_step.append((v, dt))
    value = float(np.mean([v for v, _ in per_step]))
    ms = float(np.mean([d for _, d in per_step])) * 1e3
    line = {
        'impl': 'reference', 'metric': 'dmc_walker_steps_per_sec',
        'value': value, 'unit': 'walker-steps/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args, 1),
        'cpu_baseline': {'value': value, 'unit': 'walker-steps/s',
                         'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'walker-steps/s',
                'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, world):
    return {
        'workload': 'mrbp_qmc DMC N=100 (BASELINE configs[3] shard): '
                    'V0=5pi^2 g=2 L=100 r_m=25 dt=6.25e-4 kappa=0.5',
        'boson_number': NOP,
        'target_walkers_per_gpu': args.walkers,
        'capacity_per_gpu': int(args.walkers * CAP_FACTOR),
        'global_target_walkers': args.walkers * world,
        'time_steps_per_step': args.nts,
        'estimators': 'off in the timed region',
        'parallelism': f'walkers sharded over {world} GPU(s)',
        'l2_policy': 'working set (2 x %.0f MB walker buffers) exceeds the '
                     '126 MB L2' % (args.walkers * CAP_FACTOR * 16 * NOP
                                    / 1e6),
    }


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from phd_qmclib_b200 import engine, _lib

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (there is no CPU fallback '
                         'for the product arm; use --impl reference)')
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # stdout carries exactly one JSON line: at NCCL_DEBUG=VERSION (the
        # setting of the GPU boxes) NCCL prints its version banner there
        if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':
            del os.environ['NCCL_DEBUG']
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    nw, nts = args.walkers, args.nts
    cap = int(nw * CAP_FACTOR)
    spec = model_spec()
    eng = engine.Engine(spec, device=local)
    dp = eng.dmc_params(TIME_STEP, cap * world, nw * world, NWC, SEED, 0.0,
                        float(NOP), local_capacity=cap)
    if world > 1:
        eng.comm_init_torch(dist, rank, world)
    eng.dmc_init(dp, initial_confs(nw, 100 + rank),
                 global_slot_offset=rank * cap)
    stream = torch.cuda.ExternalStream(eng.stream)

    # fp64 roofline denominator, measured on this GPU, now
    peak_burst = engine.measure_fp64_peak(local)
    peak_sust = engine.measure_fp64_peak(local, sustained_seconds=1.5)

    # ---- device-resident timing ------------------------------------------
    eng.set_profiling(True)
    moved = 0
    for _ in range(args.warmup):
        eng.dmc_advance(nts)
        if world > 1:
            eng.dmc_rebalance()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record(stream)
    ws_local, kern_ms, launches = 0.0, 0.0, 0
    series = dict(energy=np.zeros(nts), weight=np.zeros(nts),
                  num_walkers=np.zeros(nts, dtype=np.uint64),
                  ref_energy=np.zeros(nts), accum_energy=np.zeros(nts))
    local_ws_list = []
    for _ in range(args.steps):
        eng.dmc_run_block(nts, out=series)
        st = eng.last_block_stats()
        kern_ms += st['step_kernel_ms']
        launches += st['launches']
        local_ws_list.append(st.get('local_walker_steps', 0))
        ws_local += float(series['num_walkers'].sum())   # GLOBAL when world>1
        if world > 1:
            # order-preserving neighbour shifts, once per block, timed
            moved += eng.dmc_rebalance()
    e1.record(stream)
    barrier(); torch.cuda.synchronize()
    t1 = time.perf_counter()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    # with several ranks the series already hold global counts
    ws_total = ws_local
    value = ws_total / (dev_ms * 1e-3)
    ms_per_step = dev_ms / args.steps
    e_per_particle = float(series['energy'][-1] / series['weight'][-1] / NOP)

    # roofline of the step kernel (this rank's launches, this rank's walkers)
    ws_rank = ws_total / world
    F = flops_per_walker_step(NOP)
    ach_tf = ws_rank * F / (kern_ms * 1e-3) / 1e12
    hbm_gbs = ws_rank * bytes_per_walker_step(NOP) / (kern_ms * 1e-3) / 1e9
    traffic, ncu_view = None, None
    prof_path = os.path.join(ROOT, 'profiles', 'step_kernel_traffic.json')
    if os.path.exists(prof_path):
        try:
            tj = json.load(open(prof_path))
            traffic = tj['dram_bytes_per_walker'] * ws_rank / (
                args.steps * nts)
            # the executed-instruction view of the same kernel, from the
            # committed ncu capture (not measured in this run)
            ncu_view = {k: tj[k] for k in (
                'fp64_pipe_active_pct', 'issue_active_pct',
                'warp_instructions_per_launch', 'registers_per_thread',
                'source') if k in tj}
        except Exception:
            traffic = None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    roofline = {
        'kernel': 'dmc_step_kernel', 'bound': 'fp64',
        'achieved': ach_tf, 'peak': peak_sust, 'unit': 'TFLOP/s',
        'frac': ach_tf / peak_sust,
        'peak_source': 'DFMA microbenchmark measured in this run, sustained '
                       '1.5 s (MEASURED_PEAKS.json has no fp64 figure; '
                       'nominal 37.2)',
        'peak_burst': peak_burst, 'frac_of_burst': ach_tf / peak_burst,
        'flop_per_walker_step': F,
        'avg_launch_ms': kern_ms / (args.steps * nts),
        'step_kernel_share_of_step': kern_ms / max_over_ranks(
            e0.elapsed_time(e1)),
        'traffic': traffic,
        'ncu': ncu_view,
        'hbm': {'achieved': hbm_gbs, 'peak': peaks.get('hbm_gbs', 6650.0),
                'unit': 'GB/s',
                'frac': hbm_gbs / peaks.get('hbm_gbs', 6650.0),
                'bytes_per_walker_step': bytes_per_walker_step(NOP)},
    }

    # ---- end to end through the C ABI with host buffers -------------------
    eng.set_profiling(False)
    nx = eng.dmc_get_next()
    n_live = int(nx['scalars'].num_walkers)
    h_confs = engine.pinned_empty((cap, 2, NOP))
    h_energy = engine.pinned_empty((cap,))
    h_weight = engine.pinned_empty((cap,))
    h_slot = engine.pinned_empty((cap,))
    h_confs[:n_live] = nx['confs']; h_energy[:n_live] = nx['energy']
    h_weight[:n_live] = nx['weight']; h_slot[:] = nx['slot_energy']
    sc = nx['scalars']
    h2d = d2h = 0
    e2e_ws = 0.0

    def e2e_step(sc, n_live):
        eng.dmc_set_state(dp, h_confs[:n_live], h_energy[:n_live],
                          h_weight[:n_live], sc, slot_energy=h_slot,
                          global_slot_offset=rank * cap)
        eng.dmc_run_block(nts, out=series)
        if world > 1:
            eng.dmc_rebalance()
        sc2 = eng.dmc_get_next_into(h_confs, h_energy, h_weight, h_slot)
        return sc2, int(sc2.num_walkers)

    sc, n_live = e2e_step(sc, n_live)        # warm-up
    barrier(); torch.cuda.synchronize()
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        n_in = n_live
        sc, n_live = e2e_step(sc, n_live)
        h2d += n_in * (2 * NOP + 2) * 8 + cap * 8
        d2h += n_live * (2 * NOP + 2) * 8 + cap * 8 + nts * 5 * 8
        e2e_ws += float(series['num_walkers'].sum())
    e1.record(stream)
    barrier(); torch.cuda.synchronize()
    e2e_wall = max_over_ranks(time.perf_counter() - t0)
    e2e = {'value': e2e_ws / e2e_wall, 'unit': 'walker-steps/s',
           'h2d_bytes_per_step': int(sum_over_ranks(h2d) / args.steps),
           'd2h_bytes_per_step': int(sum_over_ranks(d2h) / args.steps),
           'timing': 'host wall clock around K x (set_state from pinned '
                     'host + run_block + get_next to pinned host), max over '
                     'ranks',
           'ms_per_step': e2e_wall * 1e3 / args.steps}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, cores, sample, _ = cpu_port_walker_steps_per_s()
        cpu_baseline = {'value': v, 'unit': 'walker-steps/s', 'cores': cores,
                        'kind': 'port', 'sample': sample}

    if rank == 0:
        line = {
            'metric': 'dmc_walker_steps_per_sec', 'value': value,
            'unit': 'walker-steps/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(args, world),
            'per_gpu_value': value / world,
            'energy_per_particle_last_step': e_per_particle,
            'roofline': roofline, 'cpu_baseline': cpu_baseline, 'e2e': e2e,
            'gpu_launches': int(launches), 'clocks': clocks,
            'rebalanced_walkers_rank0': int(moved),
            'lib': os.path.relpath(_lib.LIB_PATH, ROOT),
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--walkers', type=int, default=125000,
                    help='target walkers per GPU')
    ap.add_argument('--nts', type=int, default=128,
                    help='DMC time steps per bench step (block)')
    ap.add_argument('--no-cpu', action='store_true',
                    help='skip the cpu_baseline leg')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.gpus != world and args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
               f'--nproc-per-node={args.gpus}', '--master-addr', '127.0.0.1',
               '--master-port', '29511', os.path.abspath(__file__)] \
            + sys.argv[1:]
        return subprocess.call(cmd)
    return run_b200(args)


if __name__ == '__main__':
    sys.exit(main())
