"""Build tuning variants of libqmcb200.so (development aid).
usage: python scripts/build_variants.py name:DEF1,DEF2 name2:DEF ..."""
import os
import sys
sys.path.insert(0, '.')
from phd_qmclib_b200 import build
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(':')
    out = os.path.join('phd_qmclib_b200', f'variant_{name}.so')
    build.build(force=True, defines=[d for d in defs.split(';') if d], out=out)
    log = os.popen(f"cuobjdump -res-usage {out} 2>/dev/null | grep -A1 dmc_step | tail -1").read().strip()
    print(name, log[:120])
