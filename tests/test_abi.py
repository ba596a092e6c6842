"""CPU tests of the boundary: libdr4sr.so loads, exports every symbol include/dr4sr.h declares, the
ctypes table covers them, argument validation works without a GPU (no compute is launched)."""
import ctypes as C
import os
import re

import pytest

from dr4sr_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(REPO, 'include', 'dr4sr.h')).read()
    return sorted(set(re.findall(r'DR4SR_API\s+[\w\s\*]+?\b(dr4sr_\w+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()


def test_header_declares_something():
    assert len(declared_symbols()) >= 16


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), f'{name} declared in include/dr4sr.h but not exported'


def test_ctypes_table_matches_header(lib):
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_abi_version(lib):
    assert lib.dr4sr_abi_version() == 1


def test_sizes_and_validation_without_gpu(lib):
    cfg = _lib.SasrecCfg(B=8, L=50, D=64, F=128, n_head=2, n_layer=2, N=100, dropout_p=0.0, ln_eps=1e-12, seed=0, step=0)
    per_layer = 3 * 64 * 64 + 3 * 64 + 64 * 64 + 64 + 128 * 64 + 128 + 64 * 128 + 64 + 4 * 64
    assert lib.dr4sr_sasrec_param_count(C.byref(cfg)) == 50 * 64 + 2 * per_layer
    assert lib.dr4sr_sasrec_workspace_bytes(C.byref(cfg)) > 0
    bad = _lib.SasrecCfg(B=8, L=50, D=96, F=128, n_head=2, n_layer=2, N=100, dropout_p=0.0, ln_eps=1e-12, seed=0, step=0)
    assert lib.dr4sr_sasrec_param_count(C.byref(bad)) == 0          # unsupported width is refused, not mis-run
    assert lib.dr4sr_prep_batch(None, None, 8, 50, 0, None, None, None, None) == -1
    assert lib.dr4sr_adam(None, None, None, None, 10, 1, 1e-3, .9, .999, 1e-8, 0., 0, None) == -1
    assert lib.dr4sr_topk_workspace_bytes(4, 1000, 100) >= 4 * 1000 * 4


def test_product_refuses_cpu_tensors(lib):
    import torch
    from dr4sr_b200 import engine
    with pytest.raises(_lib.Dr4srError):
        engine.adam_step(torch.zeros(4), torch.zeros(4), torch.zeros(4), torch.zeros(4), 1, 1e-3)
