"""SASS opcode histogram of one kernel of libqmcb200.so (cuobjdump -sass).

    python scripts/sass_histogram.py KERNEL_SUBSTRING > profiles/xxx.txt
"""
import collections
import re
import subprocess
import sys

LIB = 'phd_qmclib_b200/libqmcb200.so'


def main(pattern):
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True,
                          text=True, check=True).stdout.splitlines()
    out, name, take = [], None, False
    for ln in sass:
        m = re.search(r'Function : (\S+)', ln)
        if m:
            take = pattern in m.group(1)
            name = m.group(1) if take else name
            if take:
                out.append((m.group(1), collections.Counter(), [0]))
            continue
        if take and re.match(r'\s+/\*[0-9a-f]{4}\*/', ln):
            toks = ln.split()
            op = next(t for t in toks[1:] if not t.startswith('@'))
            op = op.rstrip(';').split('.')[0]
            out[-1][1][op] += 1
            out[-1][2][0] += 1
    for fn, cnt, tot in out:
        print(f'== {fn}: {tot[0]} SASS instructions (static)')
        fp64 = sum(cnt[k] for k in ('DFMA', 'DMUL', 'DADD', 'DSETP'))
        print(f'   fp64 pipe (DFMA/DMUL/DADD/DSETP): {fp64};  MUFU: '
              f'{cnt["MUFU"]};  LDS/STS: {cnt["LDS"]}/{cnt["STS"]};  '
              f'LDG/STG: {cnt["LDG"]}/{cnt["STG"]};  BAR: {cnt["BAR"]}')
        for op, n in cnt.most_common():
            print(f'   {op:12s} {n:6d}')


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'dmc_step_kernel')
