"""Condense one kernel launch of an .ncu-rep into the executed-instruction
view bench.py attaches to `roofline.ncu` (profiles/kernel_ncu_view.json).

    python scripts/ncu_view.py REP CONFIG KERNEL_SUBSTR UNITS_PER_LAUNCH "command"
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'profiles', 'kernel_ncu_view.json')


def main(rep, config, kernel, units, command):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, launches = rows[0], rows[2:]
    ki = head.index('Kernel Name')
    row = [r for r in launches if kernel in r[ki]][-1]

    def get(name, scale=1.0):
        if name not in head:
            return None
        v = row[head.index(name)].replace(',', '')
        try:
            return float(v) * scale
        except ValueError:
            return None
    unit_of = dict(zip(head, rows[1]))

    def bytes_of(name):
        v = get(name)
        if v is None:
            return None
        u = unit_of[name].lower()
        return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}[u]
    units = float(units)
    dram = (bytes_of('dram__bytes_read.sum') or 0) \
        + (bytes_of('dram__bytes_write.sum') or 0)
    view = {
        'kernel': row[ki].split('(')[0],
        'duration_ms_under_ncu': get('gpu__time_duration.sum') if
        unit_of['gpu__time_duration.sum'] == 'ms' else
        get('gpu__time_duration.sum', 1e-3),
        'fp64_pipe_active_pct': get(
            'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'),
        'issue_active_pct': get(
            'smsp__issue_active.avg.pct_of_peak_sustained_active'),
        'warp_instructions_per_launch': get('smsp__inst_executed.sum'),
        'registers_per_thread': get('launch__registers_per_thread'),
        'dram_bytes_per_launch': dram,
        'units_per_launch': units,
        'dram_bytes_per_walker': dram / units if units else None,
        'thread_dfma': get('smsp__sass_thread_inst_executed_op_dfma_pred_on.sum'),
        'thread_dadd': get('smsp__sass_thread_inst_executed_op_dadd_pred_on.sum'),
        'thread_dmul': get('smsp__sass_thread_inst_executed_op_dmul_pred_on.sum'),
        'source': f'{os.path.basename(rep)} (ncu --set full --clock-control '
                  f'none, {command})',
    }
    if view['thread_dfma'] is not None and view['duration_ms_under_ncu']:
        flop = 2 * view['thread_dfma'] + (view['thread_dadd'] or 0) \
            + (view['thread_dmul'] or 0)
        view['executed_fp64_flop_per_launch'] = flop
        view['executed_tflops_under_ncu'] = flop / (
            view['duration_ms_under_ncu'] * 1e-3) / 1e12
    try:
        allv = json.load(open(OUT))
    except Exception:
        allv = {}
    allv[config] = view
    json.dump(allv, open(OUT, 'w'), indent=1)
    print(json.dumps(view, indent=1))


if __name__ == '__main__':
    main(*sys.argv[1:6])
