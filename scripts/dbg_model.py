import sys, numpy as np
sys.path.insert(0, '.')
from phd_qmclib_b200 import engine
for name in sys.argv[1:]:
    g = np.load(f'tests/golden/model_{name}.npz')
    p = g['params']
    with engine.Engine((p[:12], p[12:19], p[19:])) as eng:
        o = eng.model_eval(g['confs'])
    d = np.abs(o['drift'] - g['drift'])
    for b in range(d.shape[0]):
        i = int(np.argmax(d[b]))
        print(name, b, f'dE={abs(o["energy"][b]-g["energy"][b]):.2e} dln={abs(o["lnpsi"][b]-g["lnpsi"][b]):.2e} '
              f'dF={d[b].max():.3e} at i={i} z={g["confs"][b,0,i]!r} F={o["drift"][b,i]:.6f} ref={g["drift"][b,i]:.6f} maxF={np.abs(g["drift"][b]).max():.3f}')
