"""Strict relative error of lnPsi / E_L (engine vs goldens and vs oracle):
prints max |a-b|/|b| and the entry where it happens (development probe)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import golden, golden_names
import oracle
from phd_qmclib_b200 import engine

def rep(tag, got, want):
    r = np.abs(got - want) / np.abs(want)
    i = int(np.argmax(r))
    print(f'{tag:40s} max_rel={r[i]:.2e} at |b|={abs(want[i]):.3e} '
          f'median|b|={np.median(np.abs(want)):.3e}  n>1e-12={(r>1e-12).sum()}'
          f' n>1e-13={(r>1e-13).sum()} / {r.size}', flush=True)

for name in golden_names('model_'):
    g = golden(name); p = g['params']
    with engine.Engine((p[:12], p[12:19], p[19:])) as eng:
        o = eng.model_eval(g['confs'])
    rep(name + ' lnpsi(ref)', o['lnpsi'], g['lnpsi'])
    rep(name + ' energy(ref)', o['energy'], g['energy'])
    nop, size = int(p[3]), float(p[4])
    rng = np.random.default_rng(7)
    n = 3000 if nop <= 100 else 300
    confs = np.zeros((n, 2, nop)); confs[:, 0] = rng.random((n, nop)) * size
    ref = oracle.model_eval(p, confs)
    with engine.Engine((p[:12], p[12:19], p[19:])) as eng:
        o = eng.model_eval(confs)
    rep(name + ' lnpsi(oracle bulk)', o['lnpsi'], ref['lnpsi'])
    rep(name + ' energy(oracle bulk)', o['energy'], ref['energy'])
