"""Row-sharded item table over PEER MEMORY (NVLink / NVSwitch, one process per GPU).

No reference counterpart (the reference is single-device).  Rank r owns rows [lo_r, hi_r) of the item table, its gradient
accumulator and its Adam state.  Every rank exports its table shard and gradient shard through CUDA IPC and opens the
other ranks' handles IN ITS OWN DEVICE CONTEXT (so the mapping is a peer mapping the local kernels can dereference); the
resulting pointers go into a `dr4sr_shard_map` (include/dr4sr.h).  From then on the row exchange is inside the kernels:

  * the fused forward and the scoring kernel read E rows straight from the owner's HBM,
  * the scatter kernel `red.global.add`s gradient rows straight into the owner's accumulator,
  * Adam runs on the local shard only;

there is no all-to-all, no request planning and no host synchronisation.  Two stream-ordered barriers per step order the
ranks: gathers start after every rank's Adam (the barrier that also sums the valid-target count, after the batch preparation),
Adam starts after every rank's scatter (the barrier in front of the encoder-gradient all-reduce at the end of the backward).
Both are kernels over the same peer mappings (`dr4sr_peer_barrier`, `dr4sr_peer_allreduce`: flag exchange with system-scope
release / acquire, gradient sum read straight out of every rank's staging buffer in rank order) -- the training step contains
no NCCL call.  `collectives='nccl'` keeps the NCCL variant (all-reduce of the count / of [encoder gradient | loss]).
Consequence of the in-place accumulation: ONE backward per optimizer step (the accumulator is cleared by the Adam pass only).
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch
import torch.distributed as dist
from torch.multiprocessing.reductions import reduce_tensor

from . import _lib
from ._lib import PeerComm, ShardMap, check
from .dist import shard_rows


def _open_in_local_context(fn, args, local_index: int) -> torch.Tensor:
    """Rebuild a CUDA-IPC shared tensor with the handle opened in THIS process's current device context: torch's default
    opens it in the exporting device's context, which torch ops on that device can use but kernels launched on the local
    device cannot (measured: illegal address).  Opened locally it is an ordinary peer mapping."""
    a = list(args)
    a[6] = local_index                       # storage_device of torch.multiprocessing.reductions.rebuild_cuda_tensor
    return fn(*a)


class PeerTable:
    def __init__(self, num_rows: int, embed_dim: int, group, shard_param: torch.Tensor, shard_grad: torch.Tensor,
                 n_stage: int = 0, collectives: str = 'peer') -> None:
        self.lib = _lib.lib()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _lib.MAX_SHARDS:
            raise _lib.Dr4srError(f'peer-sharded table supports at most {_lib.MAX_SHARDS} ranks, got {self.world}')
        self.N, self.D = int(num_rows), int(embed_dim)
        self.ranges = shard_rows(self.N, self.world)
        self.lo, self.hi = self.ranges[self.rank]
        dev = shard_param.device
        if tuple(shard_param.shape) != (self.hi - self.lo, self.D) or tuple(shard_grad.shape) != tuple(shard_param.shape):
            raise _lib.Dr4srError('peer table: the local shard must hold exactly the rows this rank owns')
        local = dev.index if dev.index is not None else torch.cuda.current_device()
        # peer collectives: flag array, count slots and the staging buffer the backward writes [encoder gradient | loss] into
        self.coll = collectives == 'peer' and n_stage > 0
        self.flags = torch.zeros(_lib.MAX_SHARDS, dtype=torch.int32, device=dev)
        self.slots = torch.zeros(_lib.MAX_SHARDS, dtype=torch.int32, device=dev)
        self.stage = torch.zeros(max(int(n_stage), 4), dtype=torch.float32, device=dev)
        self.epoch = 0
        mine = (local, reduce_tensor(shard_param), reduce_tensor(shard_grad), reduce_tensor(self.flags), reduce_tensor(self.slots),
                reduce_tensor(self.stage))
        objs: List = [None] * self.world
        dist.all_gather_object(objs, mine, group=group)
        self._keep = []                      # the peer mappings live as long as this object
        self.cmap = ShardMap()
        self.comm = PeerComm()
        for r, (peer_dev, (f_t, a_t), (f_g, a_g), (f_f, a_f), (f_s, a_s), (f_b, a_b)) in enumerate(objs):
            if r == self.rank:
                t, g, fl, sl, sb = shard_param, shard_grad, self.flags, self.slots, self.stage
            else:
                check(self.lib.dr4sr_enable_peer_access(int(peer_dev)), f'peer access to device {peer_dev} (no NVLink / P2P path?)')
                t = _open_in_local_context(f_t, a_t, local)
                g = _open_in_local_context(f_g, a_g, local)
                fl = _open_in_local_context(f_f, a_f, local)
                sl = _open_in_local_context(f_s, a_s, local)
                sb = _open_in_local_context(f_b, a_b, local)
            self._keep.append((t, g, fl, sl, sb))
            self.cmap.table[r] = t.data_ptr()
            self.cmap.grad[r] = g.data_ptr()
            self.cmap.lo[r] = self.ranges[r][0]
            self.comm.flags[r], self.comm.slots[r], self.comm.stage[r] = fl.data_ptr(), sl.data_ptr(), sb.data_ptr()
        self.comm.world, self.comm.rank = self.world, self.rank
        self.cmap.lo[self.world] = self.N
        self.cmap.world, self.cmap.rank = self.world, self.rank
        self._own = (shard_param.data_ptr(), shard_grad.data_ptr())
        self._tick = torch.zeros(1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)          # the zero-fills of flags / slots / stage are done ...
        dist.barrier(group=group)            # ... and every rank has mapped every buffer before anyone proceeds

    def ref(self):
        return C.byref(self.cmap)

    def check_unmoved(self, shard_param: torch.Tensor, shard_grad: torch.Tensor) -> None:
        if (shard_param.data_ptr(), shard_grad.data_ptr()) != self._own:
            raise _lib.Dr4srError('peer table: the local shard was re-allocated after enable_peer_table (the other ranks hold '
                                  'pointers into it); call enable_peer_table again on every rank')

    def barrier(self, count: torch.Tensor = None) -> None:
        """Stream-ordered barrier over the ranks (no host sync): kernels enqueued after it start once every rank's kernels
        enqueued before it have finished.  `count` (a 1-element int32 device tensor): replaced by its sum over the ranks."""
        if not self.coll:
            dist.all_reduce(self._tick, group=self.group)
            if count is not None:
                dist.all_reduce(count, group=self.group)
            return
        self.epoch += 1
        check(self.lib.dr4sr_peer_barrier(C.byref(self.comm), self.epoch, None if count is None else count.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream), 'dr4sr_peer_barrier')

    def allreduce_stage(self, n: int, out: torch.Tensor) -> None:
        """out[:n] = sum over ranks of their staging buffers' first n floats (rank order: identical bits everywhere); also the
        barrier "every rank's kernels enqueued so far have finished"."""
        if n > self.stage.numel() or out.numel() < n or out.dtype != torch.float32 or not out.is_contiguous():
            raise _lib.Dr4srError('peer all-reduce: bad sizes')
        self.epoch += 1
        check(self.lib.dr4sr_peer_allreduce(C.byref(self.comm), self.epoch, int(n), out.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream), 'dr4sr_peer_allreduce')

    def barrier_async(self):
        """The same barrier as a Work handle: it is ordered behind everything enqueued on the current stream so far, runs on
        NCCL's stream beside whatever the caller enqueues next (rank-local kernels), and `.wait()` orders the current stream
        behind it."""
        return dist.all_reduce(self._tick, group=self.group, async_op=True)
