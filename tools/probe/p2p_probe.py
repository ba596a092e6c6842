"""Feasibility probe (2 ranks, torchrun): share a CUDA tensor between the ranks with CUDA IPC (torch.multiprocessing reductions),
enable peer access, and read / atomically update the PEER's memory from a libdr4sr kernel launched on the local device."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
from torch.multiprocessing.reductions import reduce_tensor
from dr4sr_b200 import _lib
from dr4sr_b200.engine import _p, _stream

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
lib = _lib.lib()
N, D = 1000, 128
mine = torch.full((N, D), float(rank + 1), device=dev)
mine[:, 0] = torch.arange(N, device=dev).float()
fn, args = reduce_tensor(mine)
objs = [None] * world
dist.all_gather_object(objs, (fn, args))
def open_here(f, a):
    a = list(a); a[6] = local              # open the IPC handle in THIS device's context (lazy peer mapping)
    return f(*a)
peers = []
for r, (f, a) in enumerate(objs):
    peers.append(mine if r == rank else open_here(f, a))
print(rank, 'peer tensors on', [str(t.device) for t in peers], 'can_access', [torch.cuda.can_device_access_peer(local, r) for r in range(world) if r != local], flush=True)
other = (rank + 1) % world
# enable peer access local -> other (torch does it lazily for copies; make sure by a tiny p2p copy)
for r in range(world):
    _lib.check(lib.dr4sr_enable_peer_access(r), 'peer access')
ids = torch.arange(0, N, 7, device=dev, dtype=torch.int64)
out = torch.zeros(ids.numel(), D, device=dev)
_lib.check(lib.dr4sr_gather_rows(_p(peers[other]), _p(ids), 0, ids.numel(), D, _p(out), _stream()), 'gather from peer')
torch.cuda.synchronize()
ok = bool((out[:, 1] == float(other + 1)).all()) and bool((out[:, 0] == ids.float()).all())
print(rank, 'P2P gather from rank', other, 'ok =', ok, flush=True)
# remote atomics: scatter-add ones into the peer's rows
dist.barrier()
rows = torch.ones(ids.numel(), D, device=dev)
t0 = time.perf_counter()
for _ in range(20):
    _lib.check(lib.dr4sr_scatter_add_rows(_p(peers[other]), _p(ids), 0, ids.numel(), D, _p(rows), _stream()), 'scatter to peer')
torch.cuda.synchronize(); dist.barrier()
want = float(rank + 1) + 20.0
print(rank, 'remote red.add ok =', bool((mine[ids, 1] == want).all()), 'value', float(mine[0, 1]), flush=True)
# bandwidth: gather 64K rows (32 MB) from the peer vs locally
big = torch.randn(200_000, D, device=dev)
fn, args = reduce_tensor(big); objs = [None] * world; dist.all_gather_object(objs, (fn, args))
pbig = [big if r == rank else open_here(f, a) for r, (f, a) in enumerate(objs)]
idb = torch.randint(0, 200_000, (65536,), device=dev)
outb = torch.empty(65536, D, device=dev)
for name, src in (('local', big), ('peer', pbig[other])):
    for _ in range(3):
        lib.dr4sr_gather_rows(_p(src), _p(idb), 0, idb.numel(), D, _p(outb), _stream())
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        lib.dr4sr_gather_rows(_p(src), _p(idb), 0, idb.numel(), D, _p(outb), _stream())
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(rank, f'gather 65536 rows x 512 B {name}: {ms * 1e3:.1f} us = {65536 * 512 / ms / 1e6:.0f} GB/s', flush=True)
for name, dst in (('local', big), ('peer', pbig[other])):
    for _ in range(3):
        lib.dr4sr_scatter_add_rows(_p(dst), _p(idb), 0, idb.numel(), D, _p(outb), _stream())
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        lib.dr4sr_scatter_add_rows(_p(dst), _p(idb), 0, idb.numel(), D, _p(outb), _stream())
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(rank, f'red.add 65536 rows x 512 B {name}: {ms * 1e3:.1f} us = {65536 * 512 / ms / 1e6:.0f} GB/s', flush=True)
dist.barrier()
dist.destroy_process_group()
