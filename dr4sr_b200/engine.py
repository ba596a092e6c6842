"""Device-side engine: owns the workspaces and drives libdr4sr through its C ABI.

PyTorch is used for device memory and streams only; every arithmetic step of the hot path is a
kernel of libdr4sr launched on torch's current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import FmlpCfg, GruCfg, SasrecCfg, check


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, dtype: torch.dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.Dr4srError(f'{name} must be a CUDA tensor (no CPU fallback); got device {t.device}')
    if t.dtype != dtype:
        raise _lib.Dr4srError(f'{name} must be {dtype}, got {t.dtype}')
    return t if t.is_contiguous() else t.contiguous()


@dataclass
class _Buffers:
    """Everything sized by the batch size B (cached per B)."""
    tok_off: torch.Tensor
    row_seq: torch.Tensor
    counts: torch.Tensor
    ws: torch.Tensor
    q_packed: torch.Tensor
    dq: torch.Tensor
    dx0: torch.Tensor
    dscore: torch.Tensor
    loss_pos: torch.Tensor
    loss: torch.Tensor
    q_last: torch.Tensor
    fwd_train: bool = True          # whether the forward whose activations `ws` holds drew dropout masks
    fwd_step: int = 0               # ... and with which dropout-stream step


class SASRecEngine:
    """Kernels + workspaces of one SASRec encoder (reference model/sasrec.py:10-75)."""
    loss_kind = 0                       # DR4SR_LOSS_BCE; BaseModel._init_model sets it from config['model']['loss_fn']
    deterministic_scatter = False       # config['train']['deterministic_scatter']: sorted warp-segmented reduction instead of atomics

    def __init__(self, num_items: int, embed_dim: int, max_seq_len: int, hidden_size: int, head_num: int,
                 layer_num: int, dropout_rate: float, layer_norm_eps: float, seed: int, device: torch.device) -> None:
        self.lib = _lib.lib()
        self.N, self.D, self.L, self.F = int(num_items), int(embed_dim), int(max_seq_len), int(hidden_size)
        self.H, self.n_layer = int(head_num), int(layer_num)
        self.p, self.eps, self.seed = float(dropout_rate), float(layer_norm_eps), int(seed) & (2 ** 64 - 1)
        self.device = torch.device(device)
        self.step = 0                       # dropout stream counter (bumped per training forward)
        self.fwd_token = 0                  # identifies whose activations the workspace holds
        self._bufs: Dict[int, _Buffers] = {}
        self._cfgs: Dict[int, SasrecCfg] = {}
        self._fn_count, self._fn_ws = self.lib.dr4sr_sasrec_param_count, self.lib.dr4sr_sasrec_workspace_bytes
        self._fn_fwd, self._fn_bwd, self._name = self.lib.dr4sr_sasrec_fwd, self.lib.dr4sr_sasrec_bwd, 'dr4sr_sasrec'
        self._fn_bwd_async = self.lib.dr4sr_sasrec_bwd_async
        self._finish_init(f'unsupported SASRec shape D={self.D} F={self.F} L={self.L} heads={self.H} layers={self.n_layer} '
                          f'(D in {{64,128}}, F % 64 == 0, L <= 64)')

    def _finish_init(self, why: str) -> None:
        n = self._fn_count(C.byref(self.cfg(1)))
        if n == 0:
            raise _lib.Dr4srError(why)
        self.param_count = int(n)
        self._tg_ws = torch.empty(self.lib.dr4sr_table_grad_workspace_bytes(self.L, self.D), dtype=torch.uint8, device=self.device)

    # ---- configuration ------------------------------------------------------------------------
    def cfg(self, B: int, step: Optional[int] = None) -> SasrecCfg:
        c = self._cfgs.get(B) if hasattr(self, '_cfgs') else None
        if c is None:
            c = SasrecCfg(B=B, L=self.L, D=self.D, F=self.F, n_head=self.H, n_layer=self.n_layer, N=self.N,
                          dropout_p=self.p, ln_eps=self.eps, seed=self.seed, step=0)
            if hasattr(self, '_cfgs'):
                self._cfgs[B] = c
        c.step = self.step if step is None else step      # the struct is read during the (synchronous) C call only
        return c

    def buffers(self, B: int) -> _Buffers:
        b = self._bufs.get(B)
        if b is None:
            dev, T, D = self.device, B * self.L, self.D
            ws_bytes = self._fn_ws(C.byref(self.cfg(B)))
            f32 = dict(dtype=torch.float32, device=dev)
            b = _Buffers(
                tok_off=torch.zeros(B + 1, dtype=torch.int32, device=dev),
                row_seq=torch.zeros(T, dtype=torch.int32, device=dev),
                counts=torch.zeros(4, dtype=torch.int32, device=dev),
                ws=torch.empty(ws_bytes, dtype=torch.uint8, device=dev),
                q_packed=torch.zeros(T, D, **f32), dq=torch.zeros(T, D, **f32), dx0=torch.zeros(T, D, **f32),
                dscore=torch.zeros(T, 2, **f32), loss_pos=torch.zeros(B, self.L, **f32), loss=torch.zeros((), **f32),
                q_last=torch.zeros(B, D, **f32))
            self._bufs[B] = b
        return b

    # ---- kernels ------------------------------------------------------------------------------
    def prep(self, seqlen: torch.Tensor, item_id: Optional[torch.Tensor]) -> _Buffers:
        seqlen = _req(seqlen, torch.int64, 'seqlen')
        B = seqlen.numel()
        b = self.buffers(B)
        one_d = 0
        if item_id is not None:
            item_id = _req(item_id, torch.int64, 'item_id')
            one_d = 1 if item_id.dim() == 1 else 0
        check(self.lib.dr4sr_prep_batch(_p(seqlen), _p(item_id), B, self.L, one_d, _p(b.tok_off), _p(b.row_seq),
                                        _p(b.counts), _stream()), 'dr4sr_prep_batch')
        return b

    def encode(self, b: _Buffers, table: torch.Tensor, flat: torch.Tensor, in_ids: torch.Tensor, train: bool,
               want_last: bool = False, q_dense: Optional[torch.Tensor] = None) -> torch.Tensor:
        in_ids = _req(in_ids, torch.int64, 'in_item_id')
        B = in_ids.size(0)
        cfg = self.cfg(B)
        b.fwd_train, b.fwd_step = bool(train), self.step
        self.fwd_token += 1                 # the workspace now holds THIS forward's activations (stale backwards are refused)
        if hasattr(table, 'cmap'):          # row-sharded table over peer memory (dr4sr_b200/peer.py)
            if self._name != 'dr4sr_sasrec':
                raise _lib.Dr4srError('the peer-sharded table is implemented for SASRec; use enable_sharded_table for ' + self._name)
            check(self.lib.dr4sr_sasrec_fwd_sharded(C.byref(cfg), table.ref(), _p(_req(flat, torch.float32, 'params')), _p(in_ids),
                                                    _p(b.tok_off), _p(b.row_seq), _p(b.counts), _p(b.ws), b.ws.numel(),
                                                    1 if train else 0, _p(b.q_packed), _p(b.q_last) if want_last else None,
                                                    _p(q_dense), _stream()), 'dr4sr_sasrec_fwd_sharded')
            return b.q_packed
        check(self._fn_fwd(C.byref(cfg), _p(_req(table, torch.float32, 'table')),
                                        _p(_req(flat, torch.float32, 'params')), _p(in_ids), _p(b.tok_off), _p(b.row_seq),
                                        _p(b.counts), _p(b.ws), b.ws.numel(), 1 if train else 0, _p(b.q_packed),
                                        _p(b.q_last) if want_last else None, _p(q_dense), _stream()), self._name + '_fwd')
        return b.q_packed

    def score_bce(self, b: _Buffers, table: torch.Tensor, item_id: torch.Tensor, neg_item: torch.Tensor,
                  want_grad: bool, loss_weight: Optional[torch.Tensor] = None, upstream: Optional[torch.Tensor] = None,
                  q_packed: Optional[torch.Tensor] = None) -> torch.Tensor:
        item_id = _req(item_id, torch.int64, 'item_id')
        neg_item = _req(neg_item, torch.int64, 'neg_item')
        B = item_id.size(0)
        q = b.q_packed if q_packed is None else q_packed
        if hasattr(table, 'cmap'):
            check(self.lib.dr4sr_score_loss_sharded(self.loss_kind, _p(q), table.ref(), _p(item_id), _p(neg_item), _p(b.tok_off),
                                                    _p(b.row_seq), _p(b.counts), B, self.L, self.D, _p(loss_weight), _p(upstream),
                                                    _p(b.loss_pos), _p(b.dscore), _p(b.dq) if want_grad else None, _stream()),
                  'dr4sr_score_loss_sharded')
            return b.loss_pos
        check(self.lib.dr4sr_score_loss(self.loss_kind, _p(q), _p(table), _p(item_id), _p(neg_item), _p(b.tok_off), _p(b.row_seq), _p(b.counts),
                                       B, self.L, self.D, _p(loss_weight), _p(upstream), _p(b.loss_pos), _p(b.dscore),
                                       _p(b.dq) if want_grad else None, _stream()), 'dr4sr_score_loss')
        return b.loss_pos

    def reduce_loss(self, b: _Buffers) -> torch.Tensor:
        """Sum of the per-slot loss terms into a FRESH scalar (the caller keeps it; no clone needed)."""
        out = torch.empty((), dtype=torch.float32, device=self.device)
        check(self.lib.dr4sr_sum(_p(b.loss_pos), b.loss_pos.numel(), _p(out), _stream()), 'dr4sr_sum')
        return out

    def scale_grads(self, b: _Buffers, upstream: torch.Tensor) -> None:
        """dq, dscore (computed in the forward sweep with upstream = 1) *= upstream; a no-op kernel when it is 1."""
        check(self.lib.dr4sr_scale_grads(_p(_req(upstream, torch.float32, 'upstream')), _p(b.counts), self.D, _p(b.dq), _p(b.dscore),
                                         _stream()), 'dr4sr_scale_grads')

    def encode_bwd(self, b: _Buffers, table: torch.Tensor, flat: torch.Tensor, in_ids: torch.Tensor, grads_flat: torch.Tensor,
                   dq: Optional[torch.Tensor] = None, defer_join: bool = False) -> torch.Tensor:
        """defer_join: (SASRec only) return with the weight gradients still in flight on the library's side stream;
        the caller must call join_bwd() before anything reads `grads_flat` (dx0 is already ordered)."""
        in_ids = _req(in_ids, torch.int64, 'in_item_id')
        B = in_ids.size(0)
        cfg = self.cfg(B, step=b.fwd_step)
        if not b.fwd_train and cfg.dropout_p != 0.0:      # an eval-mode forward applied no masks: the backward must not either
            cfg = type(cfg).from_buffer_copy(cfg)
            cfg.dropout_p = 0.0
        dq = b.dq if dq is None else dq
        fn = self._fn_bwd_async if (defer_join and getattr(self, '_fn_bwd_async', None) is not None) else self._fn_bwd
        tptr = None if hasattr(table, 'cmap') else _p(table)       # (the backward never reads the table)
        check(fn(C.byref(cfg), tptr, _p(flat), _p(in_ids), _p(b.tok_off), _p(b.row_seq), _p(b.counts),
                                        _p(b.ws), b.ws.numel(), _p(dq), _p(grads_flat), _p(b.dx0), _stream()), self._name + '_bwd')
        return b.dx0

    def join_bwd(self) -> None:
        if getattr(self, '_fn_bwd_async', None) is not None:
            check(self.lib.dr4sr_sasrec_bwd_join(_stream()), 'dr4sr_sasrec_bwd_join')

    def table_grad_targets_async(self, b: _Buffers, item_id: torch.Tensor, neg_item: torch.Tensor, table_grad) -> None:
        """dE[item_id] += ds+ q, dE[neg] += ds- q on the library's background stream (runs under the encoder backward); pair with
        table_grad(..., item_id=None, neg_item=None) and table_grad_join()."""
        B = item_id.size(0)
        if hasattr(table_grad, 'cmap'):
            check(self.lib.dr4sr_table_grad_targets_async_sharded(_p(b.q_packed), _p(b.dscore), _p(item_id), _p(neg_item), _p(b.tok_off),
                                                                  _p(b.row_seq), _p(b.counts), B, self.L, self.D, table_grad.ref(),
                                                                  _stream()), 'dr4sr_table_grad_targets_async_sharded')
            return
        check(self.lib.dr4sr_table_grad_targets_async(_p(b.q_packed), _p(b.dscore), _p(item_id), _p(neg_item), _p(b.tok_off), _p(b.row_seq),
                                                      _p(b.counts), B, self.L, self.D, self.N, _p(table_grad), _stream()),
              'dr4sr_table_grad_targets_async')

    def _sorted_ws(self, B: int, rows: int) -> torch.Tensor:
        cache = self.__dict__.setdefault('_tg_sorted', {})
        ws = cache.get((B, rows))
        if ws is None:
            ws = cache[(B, rows)] = torch.empty(self.lib.dr4sr_table_grad_sorted_workspace_bytes(B, self.L, self.D, rows),
                                                dtype=torch.uint8, device=self.device)
        return ws

    def table_grad_join(self) -> None:
        check(self.lib.dr4sr_table_grad_targets_join(_stream()), 'dr4sr_table_grad_targets_join')

    def table_grad(self, b: _Buffers, in_ids: torch.Tensor, item_id: Optional[torch.Tensor], neg_item: Optional[torch.Tensor],
                   table_grad: torch.Tensor, pos_grad: Optional[torch.Tensor], with_dx0: bool = True) -> None:
        B = in_ids.size(0)
        if hasattr(table_grad, 'cmap'):     # peer-sharded accumulator: gradient rows are added in the owner's HBM
            check(self.lib.dr4sr_table_grad_sharded(_p(b.dx0) if with_dx0 else None, _p(b.q_packed), _p(b.dscore), _p(in_ids),
                                                    _p(item_id), _p(neg_item), _p(b.tok_off), _p(b.row_seq), _p(b.counts), B, self.L,
                                                    self.D, table_grad.ref(), _p(pos_grad), _p(self._tg_ws), self._tg_ws.numel(),
                                                    _stream()), 'dr4sr_table_grad_sharded')
            return
        if self.deterministic_scatter:      # sort by id + fixed-order segment sums: bit-reproducible gradients
            rows = table_grad.size(0)
            ws = self._sorted_ws(B, rows)
            check(self.lib.dr4sr_table_grad_sorted(_p(b.dx0) if with_dx0 else None, _p(b.q_packed), _p(b.dscore), _p(in_ids), _p(item_id),
                                                   _p(neg_item), _p(b.tok_off), _p(b.row_seq), _p(b.counts), B, self.L, self.D, rows,
                                                   _p(table_grad), _p(pos_grad), _p(ws), ws.numel(), _stream()), 'dr4sr_table_grad_sorted')
            return
        check(self.lib.dr4sr_table_grad(_p(b.dx0) if with_dx0 else None, _p(b.q_packed), _p(b.dscore), _p(in_ids), _p(item_id),
                                        _p(neg_item), _p(b.tok_off), _p(b.row_seq), _p(b.counts), B, self.L, self.D, self.N,
                                        _p(table_grad), _p(pos_grad), _p(self._tg_ws), self._tg_ws.numel(), _stream()),
              'dr4sr_table_grad')


class GRUEngine(SASRecEngine):
    """Kernels + workspaces of the GRU4Rec encoder (reference model/gru4rec.py:12-22); same packed-row
    machinery, scoring, scatter and buffers as the SASRec engine."""

    def __init__(self, num_items: int, embed_dim: int, max_seq_len: int, hidden_size: int, layer_num: int, dropout_rate: float,
                 seed: int, device: torch.device) -> None:
        self.lib = _lib.lib()
        self.N, self.D, self.L, self.Hh = int(num_items), int(embed_dim), int(max_seq_len), int(hidden_size)
        self.n_layer, self.p, self.seed = int(layer_num), float(dropout_rate), int(seed) & (2 ** 64 - 1)
        self.device = torch.device(device)
        self.step = 0
        self.fwd_token = 0
        self._bufs = {}
        self._fn_count, self._fn_ws = self.lib.dr4sr_gru_param_count, self.lib.dr4sr_gru_workspace_bytes
        self._fn_fwd, self._fn_bwd, self._name = self.lib.dr4sr_gru_fwd, self.lib.dr4sr_gru_bwd, 'dr4sr_gru'
        self._fn_bwd_async = None
        self._finish_init(f'unsupported GRU4Rec shape D={self.D} H={self.Hh} layers={self.n_layer} '
                          f'(D in {{64,128}}, H in {{64,128,256}}, layers <= 4)')
        self._tg_ws = torch.empty(self.lib.dr4sr_table_grad_workspace_bytes(self.L, self.D), dtype=torch.uint8, device=self.device)

    def cfg(self, B: int, step: Optional[int] = None) -> GruCfg:
        return GruCfg(B=B, L=self.L, D=self.D, H=self.Hh, n_layer=self.n_layer, N=self.N, dropout_p=self.p, seed=self.seed,
                      step=self.step if step is None else step)


@dataclass
class _FmlpBuffers:
    ws: torch.Tensor
    q_last: torch.Tensor
    dq: torch.Tensor
    dz0: torch.Tensor
    dscore: torch.Tensor
    loss_pos: torch.Tensor
    loss: torch.Tensor
    ones: torch.Tensor
    full: torch.Tensor
    tok_off1: torch.Tensor
    row_seq1: torch.Tensor
    counts: torch.Tensor          # 1-D target problem: counts[1] = number of non-pad targets
    tok_offD: torch.Tensor
    row_seqD: torch.Tensor
    countsD: torch.Tensor         # dense [B, L] problem
    q_packed: torch.Tensor = None  # alias of q_last (engine-generic name used by the data-parallel hooks)
    fwd_train: bool = True         # whether the forward whose activations `ws` holds drew dropout masks


class FMLPEngine:
    """Kernels + workspaces of the FMLP encoder (reference model/fmlp.py:8-39, module/layers.py:740-808)."""
    loss_kind = 0

    def __init__(self, num_items: int, embed_dim: int, max_seq_len: int, layer_num: int, dropout_rate: float, seed: int,
                 device: torch.device, layer_norm_eps: float = 1e-12) -> None:
        self.lib = _lib.lib()
        self.N, self.D, self.L, self.n_layer = int(num_items), int(embed_dim), int(max_seq_len), int(layer_num)
        self.p, self.eps, self.seed = float(dropout_rate), float(layer_norm_eps), int(seed) & (2 ** 64 - 1)
        self.device = torch.device(device)
        self.step = 0
        self.fwd_token = 0
        self._bufs: Dict[int, _FmlpBuffers] = {}
        n = self.lib.dr4sr_fmlp_param_count(C.byref(self.cfg(1)))
        if n == 0:
            raise _lib.Dr4srError(f'unsupported FMLP shape D={self.D} L={self.L} layers={self.n_layer} (D in {{64,128}}, L == 50)')
        self.param_count = int(n)
        self._tg_ws = torch.empty(self.lib.dr4sr_table_grad_workspace_bytes(self.L, self.D), dtype=torch.uint8, device=self.device)

    def cfg(self, B: int) -> FmlpCfg:
        return FmlpCfg(B=B, L=self.L, D=self.D, n_layer=self.n_layer, N=self.N, dropout_p=self.p, ln_eps=self.eps,
                       seed=self.seed, step=self.step)

    def buffers(self, B: int) -> _FmlpBuffers:
        b = self._bufs.get(B)
        if b is None:
            dev, T, D = self.device, B * self.L, self.D
            f32, i32 = dict(dtype=torch.float32, device=dev), dict(dtype=torch.int32, device=dev)
            ws_bytes = self.lib.dr4sr_fmlp_workspace_bytes(C.byref(self.cfg(B)))
            b = _FmlpBuffers(
                ws=torch.empty(ws_bytes, dtype=torch.uint8, device=dev), q_last=torch.zeros(B, D, **f32),
                dq=torch.zeros(B, D, **f32), dz0=torch.zeros(T, D, **f32), dscore=torch.zeros(B, 2, **f32),
                loss_pos=torch.zeros(B, **f32), loss=torch.zeros((), **f32),
                ones=torch.ones(B, dtype=torch.int64, device=dev), full=torch.full((B,), self.L, dtype=torch.int64, device=dev),
                tok_off1=torch.zeros(B + 1, **i32), row_seq1=torch.zeros(B, **i32), counts=torch.zeros(4, **i32),
                tok_offD=torch.zeros(B + 1, **i32), row_seqD=torch.zeros(T, **i32), countsD=torch.zeros(4, **i32))
            b.q_packed = b.q_last
            check(self.lib.dr4sr_prep_batch(_p(b.full), None, B, self.L, 0, _p(b.tok_offD), _p(b.row_seqD), _p(b.countsD), _stream()),
                  'dr4sr_prep_batch')
            self._bufs[B] = b
        return b

    def prep(self, item_id: Optional[torch.Tensor], B: int) -> _FmlpBuffers:
        b = self.buffers(B)
        if item_id is not None:
            item_id = _req(item_id, torch.int64, 'item_id')
        check(self.lib.dr4sr_prep_batch(_p(b.ones), _p(item_id), B, 1, 1, _p(b.tok_off1), _p(b.row_seq1), _p(b.counts), _stream()),
              'dr4sr_prep_batch')
        return b

    def encode(self, b: _FmlpBuffers, table: torch.Tensor, flat: torch.Tensor, in_ids: torch.Tensor, train: bool) -> torch.Tensor:
        in_ids = _req(in_ids, torch.int64, 'in_item_id')
        if in_ids.size(1) != self.L:
            raise _lib.Dr4srError(f'FMLP expects sequences of length {self.L}, got {in_ids.size(1)}')
        B = in_ids.size(0)
        b.fwd_train = bool(train)
        self.fwd_token += 1
        check(self.lib.dr4sr_fmlp_fwd(C.byref(self.cfg(B)), _p(_req(table, torch.float32, 'table')), _p(flat), _p(in_ids), _p(b.ws),
                                      b.ws.numel(), 1 if train else 0, _p(b.q_last), _stream()), 'dr4sr_fmlp_fwd')
        return b.q_last

    def score_bce(self, b: _FmlpBuffers, table: torch.Tensor, item_id: torch.Tensor, neg_item: torch.Tensor, want_grad: bool,
                  loss_weight: Optional[torch.Tensor] = None, upstream: Optional[torch.Tensor] = None) -> torch.Tensor:
        B = item_id.numel()
        check(self.lib.dr4sr_score_loss(self.loss_kind, _p(b.q_last), _p(table), _p(_req(item_id, torch.int64, 'item_id')),
                                       _p(_req(neg_item, torch.int64, 'neg_item')), _p(b.tok_off1), _p(b.row_seq1), _p(b.counts), B, 1,
                                       self.D, _p(loss_weight), _p(upstream), _p(b.loss_pos), _p(b.dscore),
                                       _p(b.dq) if want_grad else None, _stream()), 'dr4sr_score_loss')
        return b.loss_pos

    def reduce_loss(self, b: _FmlpBuffers) -> torch.Tensor:
        out = torch.empty((), dtype=torch.float32, device=self.device)
        check(self.lib.dr4sr_sum(_p(b.loss_pos), b.loss_pos.numel(), _p(out), _stream()), 'dr4sr_sum')
        return out

    def encode_bwd(self, b: _FmlpBuffers, table: torch.Tensor, flat: torch.Tensor, in_ids: torch.Tensor, grads_flat: torch.Tensor) -> None:
        B = in_ids.size(0)
        cfg = self.cfg(B)
        if not b.fwd_train:
            cfg.dropout_p = 0.0                            # an eval-mode forward applied no masks
        check(self.lib.dr4sr_fmlp_bwd(C.byref(cfg), _p(table), _p(flat), _p(in_ids), _p(b.ws), b.ws.numel(), _p(b.dq),
                                      _p(grads_flat), _p(b.dz0), _stream()), 'dr4sr_fmlp_bwd')

    def table_grad(self, b: _FmlpBuffers, in_ids: torch.Tensor, item_id: torch.Tensor, neg_item: torch.Tensor,
                   table_grad: torch.Tensor, pos_grad: torch.Tensor) -> None:
        B = in_ids.size(0)
        # targets / negatives: dE[item_id] += ds+ q, dE[neg] += ds- q   (1 slot per sequence)
        check(self.lib.dr4sr_table_grad(None, _p(b.q_last), _p(b.dscore), _p(item_id), _p(item_id), _p(neg_item), _p(b.tok_off1),
                                        _p(b.row_seq1), _p(b.counts), B, 1, self.D, self.N, _p(table_grad), None, None, 0, _stream()),
              'dr4sr_table_grad')
        # inputs: dE[in_id] += dz0 for every slot of the dense batch; dP[t] = sum_b dz0[b, t]
        check(self.lib.dr4sr_table_grad(_p(b.dz0), None, None, _p(in_ids), None, None, _p(b.tok_offD), _p(b.row_seqD), _p(b.countsD),
                                        B, self.L, self.D, self.N, _p(table_grad), _p(pos_grad), _p(self._tg_ws), self._tg_ws.numel(),
                                        _stream()), 'dr4sr_table_grad')


# ---- stateless wrappers ---------------------------------------------------------------------------
def neg_sample(shape, num_items: int, seed: int, step: int, device) -> torch.Tensor:
    """BaseModel._neg_sampling replacement: uniform ids in {1..N-1} (reference model/basemodel.py:50-61)."""
    out = torch.empty(shape, dtype=torch.int64, device=device)
    check(_lib.lib().dr4sr_neg_sample(_p(out), out.numel(), num_items, seed & (2 ** 64 - 1), step & (2 ** 64 - 1), _stream()),
          'dr4sr_neg_sample')
    return out


def adam_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int, lr: float, beta1: float = 0.9,
              beta2: float = 0.999, eps: float = 1e-8, weight_decay: float = 0.0, zero_grad: bool = False) -> None:
    for t, n in ((p, 'p'), (g, 'g'), (m, 'm'), (v, 'v')):
        _req(t, torch.float32, n)
        if not t.is_contiguous():
            raise _lib.Dr4srError(f'adam: {n} must be contiguous')
    check(_lib.lib().dr4sr_adam(_p(p), _p(g), _p(m), _p(v), p.numel(), int(step), lr, beta1, beta2, eps, weight_decay,
                                1 if zero_grad else 0, _stream()), 'dr4sr_adam')


def rank_metrics(topk_ids: torch.Tensor, target: torch.Tensor, cutoffs, sums: torch.Tensor) -> None:
    """sums[2i] += sum_u ndcg@cutoffs[i], sums[2i+1] += sum_u recall@cutoffs[i] (reference evaluation/__init__.py:9-36,107-134)."""
    topk_ids = _req(topk_ids, torch.int64, 'topk_ids')
    target = _req(target.reshape(-1), torch.int64, 'target')
    if not (sums.is_cuda and sums.dtype == torch.float64 and sums.numel() == 2 * len(cutoffs) and sums.is_contiguous()):
        raise _lib.Dr4srError('rank_metrics: sums must be a contiguous CUDA float64 tensor of 2 * len(cutoffs) elements')
    if target.numel() != topk_ids.size(0) or not topk_ids.is_contiguous() or not target.is_contiguous():
        raise _lib.Dr4srError('rank_metrics: one target per row of contiguous topk_ids')
    arr = (C.c_int32 * len(cutoffs))(*[int(c) for c in cutoffs])
    check(_lib.lib().dr4sr_rank_metrics(_p(topk_ids), _p(target), topk_ids.size(0), topk_ids.size(1), arr, len(cutoffs), _p(sums), _stream()),
          'dr4sr_rank_metrics')


_topk_ws: Dict[tuple, torch.Tensor] = {}


def topk(q: torch.Tensor, table: torch.Tensor, item_dead: Optional[torch.Tensor], user_hist: Optional[torch.Tensor], k: int):
    """BaseModel.topk replacement (reference model/basemodel.py:354-365)."""
    L = _lib.lib()
    q = _req(q, torch.float32, 'query')
    table = _req(table, torch.float32, 'table')
    B, D = q.shape
    N = table.size(0)
    H = 0
    if user_hist is not None:
        user_hist = _req(user_hist, torch.int64, 'user_hist')
        H = user_hist.size(1)
    if item_dead is not None:
        item_dead = _req(item_dead, torch.uint8, 'item_dead')
    nbytes = L.dr4sr_topk_workspace_bytes(B, N, k)
    key = (q.device, nbytes)
    ws = _topk_ws.get(key)
    if ws is None:
        _topk_ws.clear()
        ws = _topk_ws[key] = torch.empty(nbytes, dtype=torch.uint8, device=q.device)
    scores = torch.empty(B, k, dtype=torch.float32, device=q.device)
    ids = torch.empty(B, k, dtype=torch.int64, device=q.device)
    check(L.dr4sr_topk(_p(q), _p(table), _p(item_dead), _p(user_hist), B, D, N, H, k, _p(scores), _p(ids), _p(ws), nbytes,
                       _stream()), 'dr4sr_topk')
    return scores, ids
