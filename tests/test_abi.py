"""CPU: the C-ABI library loads and exports every symbol include/qmcb200.h
declares; creating an engine without a GPU fails loudly (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'qmcb200.h')).read()
    return sorted(set(re.findall(r'QMCB_API\s+[\w\s\*]+?\b(qmcb_\w+)\s*\(',
                                 text)))


def test_header_symbols_exported():
    from phd_qmclib_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from phd_qmclib_b200 import build
        build.build()
    L = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(L, name), f'{name} not exported'
    assert sorted(_lib.SYMBOLS) == declared
    assert b'sm_100a' in L.qmcb_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from phd_qmclib_b200 import engine
    g = golden('model_ll_n16.npz')
    p = g['params']
    with pytest.raises(engine.EngineError, match='no CUDA device'):
        engine.Engine((p[:12], p[12:19], p[19:]))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may
    import, load or execute it."""
    pkg = os.path.join(ROOT, 'phd_qmclib_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'libqmc_oracle' not in text, f
                assert 'import oracle' not in text, f
                assert 'refshim' not in text, f
