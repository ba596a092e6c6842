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


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """sizeof / offsetof of every struct of include/dr4sr.h, printed by a C program compiled with gcc against the header, equal
    the ctypes mirrors in dr4sr_b200/_lib.py (a silent mismatch would hand the kernels garbage pointers)."""
    import subprocess
    mirrors = {'dr4sr_sasrec_cfg': _lib.SasrecCfg, 'dr4sr_gru_cfg': _lib.GruCfg, 'dr4sr_fmlp_cfg': _lib.FmlpCfg,
               'dr4sr_shard_map': _lib.ShardMap, 'dr4sr_peer_comm': _lib.PeerComm}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "dr4sr.h"', 'int main(void) {']
    for cname, cls in mirrors.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-I', os.path.join(REPO, 'include'), str(src), '-o', str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    seen = 0
    for ln in out.strip().splitlines():
        cname, what, val = ln.split()
        cls = mirrors[cname]
        want = C.sizeof(cls) if what == 'size' else getattr(cls, what).offset
        assert int(val) == want, f'{cname}.{what}: header {val} vs ctypes {want}'
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in mirrors.values())


def test_new_entry_points_validate_without_gpu(lib):
    comm = _lib.PeerComm()
    comm.world, comm.rank = 2, 0                                  # null flag pointers: refused before any launch
    assert lib.dr4sr_peer_barrier(C.byref(comm), 1, None, None) == -1
    assert lib.dr4sr_peer_allreduce(C.byref(comm), 1, 16, None, None) == -1
    assert lib.dr4sr_table_grad_sorted(None, None, None, None, None, None, None, None, None, 4, 50, 128, 100, None, None, None, 0, None) == -1
    assert lib.dr4sr_table_grad_targets_async(None, None, None, None, None, None, None, 4, 50, 128, 100, None, None) == -1
    assert lib.dr4sr_rank_metrics(None, None, 4, 100, None, 1, None, None) == -1
    sm = _lib.ShardMap()
    sm.world, sm.rank = 9, 0                                      # more ranks than DR4SR_MAX_SHARDS
    assert lib.dr4sr_table_grad_targets_async_sharded(None, None, None, None, None, None, None, 4, 50, 128, C.byref(sm), None) == -1
