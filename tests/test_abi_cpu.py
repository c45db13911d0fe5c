"""CPU: the C-ABI library loads and exports every symbol include/escgnn_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'escgnn_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(escgnn_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from esc_gnn_b200 import _lib, build
    build.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(L, n), n
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert _lib.lib().escgnn_version() >= 100


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'esc_gnn_b200')
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dp, f)).read()
                assert 'oracle' not in src.replace('no CPU fallback', ''), os.path.join(dp, f)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from esc_gnn_b200.data import Data
    from esc_gnn_b200.transform import create_subgraphs
    d = Data(x=torch.ones(3, 1), edge_index=torch.tensor([[0, 1], [1, 0]]))
    with pytest.raises(RuntimeError):
        create_subgraphs(d, 2)
