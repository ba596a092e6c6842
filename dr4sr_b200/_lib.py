"""ctypes binding of libdr4sr.so (the C ABI declared in include/dr4sr.h).

The product has no CPU or PyTorch fallback: if the library is missing or cannot be loaded this
module raises at first use, loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
LIB_PATH = os.environ.get('DR4SR_LIB_PATH') or os.path.join(CSRC, 'libdr4sr.so')   # override: debug builds only (make libdr4sr_trace.so)

c_i32, c_i64, c_u64, c_f32, c_sz, c_p = C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_size_t, C.c_void_p


class FmlpCfg(C.Structure):
    """Mirror of dr4sr_fmlp_cfg."""
    _fields_ = [('B', c_i32), ('L', c_i32), ('D', c_i32), ('n_layer', c_i32), ('N', c_i64), ('dropout_p', c_f32),
                ('ln_eps', c_f32), ('seed', c_u64), ('step', c_u64)]


class GruCfg(C.Structure):
    """Mirror of dr4sr_gru_cfg."""
    _fields_ = [('B', c_i32), ('L', c_i32), ('D', c_i32), ('H', c_i32), ('n_layer', c_i32), ('N', c_i64), ('dropout_p', c_f32),
                ('seed', c_u64), ('step', c_u64)]


MAX_SHARDS = 8


class ShardMap(C.Structure):
    """Mirror of dr4sr_shard_map (row-sharded item table over peer memory)."""
    _fields_ = [('table', c_p * MAX_SHARDS), ('grad', c_p * MAX_SHARDS), ('lo', c_i64 * (MAX_SHARDS + 1)), ('world', c_i32),
                ('rank', c_i32)]


class PeerComm(C.Structure):
    """Mirror of dr4sr_peer_comm (flag arrays, count slots and staging buffers of every rank, as local / peer pointers)."""
    _fields_ = [('flags', c_p * MAX_SHARDS), ('slots', c_p * MAX_SHARDS), ('stage', c_p * MAX_SHARDS), ('world', c_i32), ('rank', c_i32)]


class SasrecCfg(C.Structure):
    """Mirror of dr4sr_sasrec_cfg."""
    _fields_ = [('B', c_i32), ('L', c_i32), ('D', c_i32), ('F', c_i32), ('n_head', c_i32), ('n_layer', c_i32),
                ('N', c_i64), ('dropout_p', c_f32), ('ln_eps', c_f32), ('seed', c_u64), ('step', c_u64)]


# name -> (restype, argtypes); must list every symbol include/dr4sr.h declares (tests check this)
SIGNATURES = {
    'dr4sr_abi_version': (c_i32, []),
    'dr4sr_last_cuda_error': (C.c_char_p, []),
    'dr4sr_enable_peer_access': (c_i32, [c_i32]),
    'dr4sr_launch_count': (C.c_longlong, []),
    'dr4sr_prof_enable': (c_i32, [c_i32]),
    'dr4sr_prof_collect': (c_sz, [C.c_char_p, c_sz]),
    'dr4sr_prep_batch': (c_i32, [c_p, c_p, c_i32, c_i32, c_i32, c_p, c_p, c_p, c_p]),
    'dr4sr_gather_i64': (c_i32, [c_p, c_i32, c_p, c_i64, c_p, c_p]),
    'dr4sr_neg_sample': (c_i32, [c_p, c_i64, c_i64, c_u64, c_u64, c_p]),
    'dr4sr_set_gemm_backend': (c_i32, [c_i32]),
    'dr4sr_set_attn_backend': (c_i32, [c_i32]),
    'dr4sr_set_fused_backend': (c_i32, [c_i32]),
    'dr4sr_fused_tiles': (c_i32, [c_p, c_i32, c_i32, c_p, c_sz, c_p]),
    'dr4sr_debug_trace': (c_i32, [C.c_void_p]),
    'dr4sr_sasrec_param_count': (c_sz, [C.POINTER(SasrecCfg)]),
    'dr4sr_sasrec_workspace_bytes': (c_sz, [C.POINTER(SasrecCfg)]),
    'dr4sr_sasrec_fwd': (c_i32, [C.POINTER(SasrecCfg), c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_sz, c_i32, c_p, c_p, c_p, c_p]),
    'dr4sr_sasrec_fwd_sharded': (c_i32, [C.POINTER(SasrecCfg), C.POINTER(ShardMap), c_p, c_p, c_p, c_p, c_p, c_p, c_sz, c_i32, c_p, c_p, c_p, c_p]),
    'dr4sr_score_loss_sharded': (c_i32, [c_i32, c_p, C.POINTER(ShardMap), c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_p, c_p, c_p, c_p, c_p, c_p]),
    'dr4sr_table_grad_sharded': (c_i32, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, C.POINTER(ShardMap), c_p, c_p, c_sz, c_p]),
    'dr4sr_sasrec_bwd': (c_i32, [C.POINTER(SasrecCfg), c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_sz, c_p, c_p, c_p, c_p]),
    'dr4sr_sasrec_bwd_async': (c_i32, [C.POINTER(SasrecCfg), c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_sz, c_p, c_p, c_p, c_p]),
    'dr4sr_sasrec_bwd_join': (c_i32, [c_p]),
    'dr4sr_scale_grads': (c_i32, [c_p, c_p, c_i32, c_p, c_p, c_p]),
    'dr4sr_unpack_rows': (c_i32, [c_p, c_p, c_i32, c_i32, c_i32, c_p, c_p, c_p]),
    'dr4sr_gru_param_count': (c_sz, [C.POINTER(GruCfg)]),
    'dr4sr_gru_workspace_bytes': (c_sz, [C.POINTER(GruCfg)]),
    'dr4sr_gru_fwd': (c_i32, [C.POINTER(GruCfg), c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_sz, c_i32, c_p, c_p, c_p, c_p]),
    'dr4sr_gru_bwd': (c_i32, [C.POINTER(GruCfg), c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_sz, c_p, c_p, c_p, c_p]),
    'dr4sr_fmlp_param_count': (c_sz, [C.POINTER(FmlpCfg)]),
    'dr4sr_fmlp_workspace_bytes': (c_sz, [C.POINTER(FmlpCfg)]),
    'dr4sr_fmlp_fwd': (c_i32, [C.POINTER(FmlpCfg), c_p, c_p, c_p, c_p, c_sz, c_i32, c_p, c_p]),
    'dr4sr_fmlp_bwd': (c_i32, [C.POINTER(FmlpCfg), c_p, c_p, c_p, c_p, c_sz, c_p, c_p, c_p, c_p]),
    'dr4sr_score_bce': (c_i32, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_p, c_p, c_p, c_p, c_p, c_p]),
    'dr4sr_score_loss': (c_i32, [c_i32, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_p, c_p, c_p, c_p, c_p, c_p]),
    'dr4sr_sum': (c_i32, [c_p, c_i64, c_p, c_p]),
    'dr4sr_table_grad_workspace_bytes': (c_sz, [c_i32, c_i32]),
    'dr4sr_table_grad_targets_async': (c_i32, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i64, c_p, c_p]),
    'dr4sr_table_grad_targets_async_sharded': (c_i32, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, C.POINTER(ShardMap), c_p]),
    'dr4sr_table_grad_targets_join': (c_i32, [c_p]),
    'dr4sr_peer_barrier': (c_i32, [C.POINTER(PeerComm), c_i32, c_p, c_p]),
    'dr4sr_peer_allreduce': (c_i32, [C.POINTER(PeerComm), c_i32, c_i64, c_p, c_p]),
    'dr4sr_rank_metrics': (c_i32, [c_p, c_p, c_i32, c_i32, C.POINTER(c_i32), c_i32, c_p, c_p]),
    'dr4sr_table_grad_sorted_workspace_bytes': (c_sz, [c_i32, c_i32, c_i32, c_i64]),
    'dr4sr_table_grad_sorted': (c_i32, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i64, c_p, c_p, c_p, c_sz, c_p]),
    'dr4sr_table_grad': (c_i32, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i64, c_p, c_p, c_p, c_sz, c_p]),
    'dr4sr_shard_plan': (c_i32, [c_p, c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i64, c_i32, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    'dr4sr_gather_rows': (c_i32, [c_p, c_p, c_i64, c_i64, c_i32, c_p, c_p]),
    'dr4sr_scatter_add_rows': (c_i32, [c_p, c_p, c_i64, c_i64, c_i32, c_p, c_p]),
    'dr4sr_adam': (c_i32, [c_p, c_p, c_p, c_p, c_i64, c_i64, c_f32, c_f32, c_f32, c_f32, c_f32, c_i32, c_p]),
    'dr4sr_topk_workspace_bytes': (c_sz, [c_i32, c_i64, c_i32]),
    'dr4sr_topk': (c_i32, [c_p, c_p, c_p, c_p, c_i32, c_i32, c_i64, c_i32, c_i32, c_p, c_p, c_p, c_sz, c_p]),
    'dr4sr_embed_fwd': (c_i32, [c_p, c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_f32, c_u64, c_u64, c_p, c_p]),
    'dr4sr_linear_fwd': (c_i32, [c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_p, c_p]),
}

_ERRORS = {-1: 'DR4SR_EINVAL (unsupported shape or null pointer)', -2: 'DR4SR_EWORKSPACE (workspace too small)',
           -3: 'DR4SR_ECUDA'}

_lock = threading.Lock()
_lib = None


class Dr4srError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile libdr4sr.so in-tree with nvcc for sm_100a (dr4sr_b200/csrc/Makefile)."""
    res = subprocess.run(['make', '-C', CSRC, '-j8'], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:], res.stderr[-4000:])
    if res.returncode != 0:
        raise Dr4srError(f'building libdr4sr.so failed (exit {res.returncode})')
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise Dr4srError(f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; '
                                     f'g.build()"` (nvcc, sm_100a). There is no CPU fallback.')
                handle = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype, fn.argtypes = res, args
                if handle.dr4sr_abi_version() != 1:
                    raise Dr4srError('libdr4sr.so ABI version mismatch; rebuild')
                _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        detail = _ERRORS.get(rc, str(rc))
        if rc == -3:
            detail += ': ' + lib().dr4sr_last_cuda_error().decode()
        raise Dr4srError(f'{what} failed: {detail}')
