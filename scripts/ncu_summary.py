"""Condense an .ncu-rep (read here with `ncu -i ... --page raw --csv`) into
the small per-launch table committed under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/x_ncu_raw.csv
"""
import csv
import subprocess
import sys

METRICS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'launch__shared_mem_per_block_dynamic',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__cycles_elapsed.avg.per_second',
    'smsp__warps_eligible.avg.per_cycle_active',
]


def main(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units, launches = rows[0], rows[1], rows[2:]
    ki = head.index('Kernel Name')
    names = [r[ki].split('(')[0].split('::')[-1] for r in launches]
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['metric', 'unit'] + [f'{n}#{i}' for i, n in enumerate(names)])
        for m in METRICS:
            if m not in head:
                continue
            i = head.index(m)
            w.writerow([m, units[i]] + [r[i] for r in launches])
    print(out, names)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
