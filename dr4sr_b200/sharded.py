"""Row-sharded item table over the GPUs of one box (NCCL all-to-all over NVLink/NVSwitch).

No reference counterpart (the reference is single-device).  Rank r owns rows [lo, hi) of the table, its
Adam state and its gradient; the encoder is replicated.  Per step:
  1. `dr4sr_shard_plan` buckets the row requests of the local sequences by owner (device side);
  2. the per-owner counts are exchanged (all-to-all of `world` integers; the one host sync of the step, it
     sizes the two variable all-to-alls);
  3. ids travel to the owners, which gather their rows (`dr4sr_gather_rows`) and send them back into the
     staged local table `loc` (row 0 = pad); forward / loss / backward kernels run unchanged on
     (loc, remapped ids);
  4. the rows of the local gradient table go back the same way and the owners scatter-add them into the
     gradient of their shard (`dr4sr_scatter_add_rows`); Adam runs on the shard only.
Encoder gradients and the valid-target count are all-reduced (`BaseModel._dp_sum`).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check
from .dist import shard_rows
from .engine import _p, _stream


class ShardedTable:
    def __init__(self, num_rows: int, embed_dim: int, group, device: torch.device) -> None:
        self.lib = _lib.lib()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.N, self.D, self.device = int(num_rows), int(embed_dim), torch.device(device)
        self.lo, self.hi = shard_rows(self.N, self.world)[self.rank]
        self._cap = 0
        self._state = None

    # ---- buffers sized by the request capacity 3*B*L -------------------------------------------------
    def _ensure(self, B: int, L: int) -> None:
        cap = 3 * B * L
        if cap <= self._cap:
            return
        dev = self.device
        self._cap = cap
        self.send_counts = torch.zeros(self.world, dtype=torch.int32, device=dev)
        self.scratch = torch.zeros(self.world + 1, dtype=torch.int32, device=dev)
        self.send_ids = torch.zeros(cap, dtype=torch.int64, device=dev)
        self.loc = torch.zeros(1 + cap, self.D, dtype=torch.float32, device=dev)
        self.gloc = torch.zeros(1 + cap, self.D, dtype=torch.float32, device=dev)
        # everything the step needs is allocated once: remapped ids, the id / row / gradient exchange buffers (their used
        # prefixes change per step, their capacity does not: a rank receives at most `world` x cap requests in the worst
        # case, sized lazily on first overflow), the count exchange
        self._loc_ids = torch.zeros(3, B * L, dtype=torch.int64, device=dev)
        self.counts2 = torch.zeros(2, self.world, dtype=torch.int64, device=dev)
        self.counts_host = torch.zeros(2, self.world, dtype=torch.int64).pin_memory()
        self._recv_cap = 0
        self.bytes_last_step = 0

    def _ensure_recv(self, n_recv: int) -> None:
        if n_recv <= self._recv_cap:
            return
        self._recv_cap = int(n_recv * 1.25) + 1024
        dev = self.device
        self.recv_ids = torch.zeros(self._recv_cap, dtype=torch.int64, device=dev)
        self.rows_out = torch.zeros(self._recv_cap, self.D, dtype=torch.float32, device=dev)
        self.grad_in = torch.zeros(self._recv_cap, self.D, dtype=torch.float32, device=dev)

    def fetch(self, shard: torch.Tensor, bufs, in_ids: torch.Tensor, item_id: Optional[torch.Tensor], neg: Optional[torch.Tensor]
              ) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]:
        """-> (local table, in_loc, item_loc, neg_loc).  `shard` = this rank's rows [lo, hi) of E."""
        B, L = in_ids.shape
        self._ensure(B, L)
        in_loc = self._loc_ids[0, :B * L].view(B, L)
        item_loc = self._loc_ids[1, :B * L].view(B, L) if item_id is not None else None
        neg_loc = self._loc_ids[2, :B * L].view(B, L) if item_id is not None else None
        check(self.lib.dr4sr_shard_plan(_p(in_ids), _p(item_id), _p(neg), _p(bufs.tok_off), _p(bufs.row_seq), _p(bufs.counts), B, L,
                                        self.N, self.world, _p(self.send_counts), _p(self.scratch), _p(self.send_ids), _p(in_loc),
                                        _p(item_loc), _p(neg_loc), _stream()), 'dr4sr_shard_plan')
        self.counts2[0].copy_(self.send_counts)
        dist.all_to_all_single(self.counts2[1], self.counts2[0], group=self.group)
        self.counts_host.copy_(self.counts2, non_blocking=True)
        torch.cuda.current_stream().synchronize()              # the step's one host sync (sizes the two variable all-to-alls)
        sc, rc = self.counts_host[0].tolist(), self.counts_host[1].tolist()
        n_send, n_recv = sum(sc), sum(rc)
        self._ensure_recv(n_recv)
        recv_ids = self.recv_ids[:n_recv]
        dist.all_to_all_single(recv_ids, self.send_ids[:n_send], rc, sc, group=self.group)
        rows_out = self.rows_out[:n_recv]
        check(self.lib.dr4sr_gather_rows(_p(shard), _p(recv_ids), self.lo, n_recv, self.D, _p(rows_out), _stream()), 'dr4sr_gather_rows')
        dist.all_to_all_single(self.loc[1:1 + n_send], rows_out, sc, rc, group=self.group)
        self._state = (sc, rc, recv_ids, n_send, n_recv)
        # all-to-all payload this rank sends + receives per step: ids out, rows back, gradient rows out (for the bench line)
        self.bytes_last_step = (n_send + n_recv) * (8 + 2 * 4 * self.D)
        return self.loc, in_loc, item_loc, neg_loc

    def local_grad(self) -> torch.Tensor:
        """Zeroed gradient table matching `loc` (only the rows in use are cleared)."""
        n_send = self._state[3]
        self.gloc[:1 + n_send].zero_()
        return self.gloc

    def push_grads(self, shard_grad: torch.Tensor) -> None:
        """Send the rows of the local gradient table to their owners and accumulate them into `shard_grad`."""
        sc, rc, recv_ids, n_send, n_recv = self._state
        grad_in = self.grad_in[:n_recv]
        dist.all_to_all_single(grad_in, self.gloc[1:1 + n_send], rc, sc, group=self.group)
        check(self.lib.dr4sr_scatter_add_rows(_p(shard_grad), _p(recv_ids), self.lo, n_recv, self.D, _p(grad_in), _stream()),
              'dr4sr_scatter_add_rows')

    # ---- eval: top-k of every rank's queries over every shard ------------------------------------------
    def topk(self, engine_topk, query: torch.Tensor, shard: torch.Tensor, dead_shard: torch.Tensor, user_hist: Optional[torch.Tensor],
             k: int):
        B, D = query.shape
        q_all = torch.empty(self.world * B, D, dtype=query.dtype, device=query.device)
        dist.all_gather_into_tensor(q_all, query.contiguous(), group=self.group)
        h_all = None
        if user_hist is not None:
            h_all = torch.empty(self.world * B, user_hist.size(1), dtype=torch.int64, device=query.device)
            dist.all_gather_into_tensor(h_all, user_hist.contiguous(), group=self.group)
            h_all = h_all - self.lo                      # ids of other shards fall outside [0, n_local) and are ignored by the kernel
        kk = min(k, shard.size(0))
        s_loc, i_loc = engine_topk(q_all, shard, dead_shard, h_all, kk)
        i_loc = i_loc + self.lo
        s_in = torch.empty(self.world * B, kk, dtype=s_loc.dtype, device=query.device)
        i_in = torch.empty(self.world * B, kk, dtype=torch.int64, device=query.device)
        dist.all_to_all_single(s_in, s_loc, group=self.group)          # chunk r of the output = rank r's candidates for MY queries
        dist.all_to_all_single(i_in, i_loc, group=self.group)
        cand_s = s_in.view(self.world, B, kk).permute(1, 0, 2).reshape(B, self.world * kk)
        cand_i = i_in.view(self.world, B, kk).permute(1, 0, 2).reshape(B, self.world * kk)
        # merge of world*k candidates per user: a few hundred values, index plumbing (not on the training path)
        top_s, pos = torch.topk(cand_s, k, dim=1)
        return top_s, cand_i.gather(1, pos)
