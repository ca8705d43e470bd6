"""Per-source-line share of samples and instructions of one kernel in an
.ncu-rep captured with --import-source on (development aid).

    python scripts/ncu_source_regions.py REP KERNEL_REGEX [TOP]
"""
import csv
import subprocess
import sys


def main(rep, kernel, top=30):
    raw = subprocess.run(
        ['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source',
         'sass,cuda', '-k', 'regex:' + kernel],
        capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    cur, hdr, agg = None, None, {}
    for r in rows:
        if r and r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r and r[0] == 'Line No':
            hdr = r
            continue
        if len(r) > 8 and r[2] == '-' and r[0].isdigit():
            try:
                agg[(cur, int(r[0]))] = (int(r[6]), int(r[7]), r[1], r)
            except ValueError:
                pass
    tots = sum(v[0] for v in agg.values())
    tot = sum(v[1] for v in agg.values())
    print(f'samples {tots}  warp instructions {tot}')
    st = [i for i, h in enumerate(hdr)
          if h.startswith('stall_') and 'Not Issued' not in h]
    S = {hdr[i]: 0 for i in st}
    for v in agg.values():
        for i in st:
            try:
                S[hdr[i]] += int(v[3][i])
            except ValueError:
                pass
    print({k: round(100 * v / tots, 1)
           for k, v in sorted(S.items(), key=lambda kv: -kv[1])
           if v > 0.01 * tots})
    for (f, l), (s, n, src, _) in sorted(agg.items(),
                                         key=lambda kv: -kv[1][0])[:top]:
        print(f'{f}:{l:4d} samp {100 * s / tots:4.1f}% inst '
              f'{100 * n / tot:4.1f}%  {src[:88]}')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
