import os, sys
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
fn, args = reduce_tensor(mine)
objs = [None] * world
dist.all_gather_object(objs, (fn, args))
other = (rank + 1) % world
a = list(objs[other][1]); a[6] = local      # open the IPC handle in THIS device's context (lazy peer mapping)
peer = objs[other][0](*a)
print(rank, 'peer device', peer.device, 'ptr', hex(peer.data_ptr()), 'mine ptr', hex(mine.data_ptr()), flush=True)
print(rank, 'sum via local-device kernel', float(peer.sum()), flush=True)
print(rank, 'enable rc', [lib.dr4sr_enable_peer_access(r) for r in range(world)], flush=True)
ids = torch.arange(0, 8, device=dev, dtype=torch.int64)
out = torch.zeros(8, D, device=dev)
rc = lib.dr4sr_gather_rows(_p(peer), _p(ids), 0, 8, D, _p(out), _stream())
try:
    torch.cuda.synchronize()
    print(rank, 'gather rc', rc, out[:, 1].tolist(), flush=True)
except Exception as e:
    print(rank, 'gather failed', str(e)[:100], flush=True)
dist.destroy_process_group()
