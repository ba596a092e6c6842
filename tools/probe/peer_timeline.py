"""2-rank probe: where does the peer-table step spend its time?  CUDA events around every all-reduce and the step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
import bench
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local); dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
mode = sys.argv[1] if len(sys.argv) > 1 else 'peer'
c = dict(bench.CFG2); c['adam_rows'] = c['num_items'] // world
model, layout = bench.build_model('SASRec', c, dev, rank, world, sharded=mode != 'replicated')
if mode == 'peer': model.enable_peer_table(dist.group.WORLD)
elif mode == 'a2a': model.enable_sharded_table(dist.group.WORLD)
else: model.enable_data_parallel(dist.group.WORLD)
model.train()
loop = bench.StepLoop(model, layout, c, dev, rank)
orig = dist.all_reduce
marks = []
def wrapped(t, *a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = orig(t, *a, **k)
    if k.get('async_op'):
        class W:
            def wait(self_inner):
                r.wait(); e1.record(); marks.append((t.numel(), e0, e1))
        return W()
    e1.record(); marks.append((t.numel(), e0, e1))
    return r
for i in range(10): loop.resident_step(i)
torch.cuda.synchronize(); dist.barrier()
dist.all_reduce = wrapped
import torch.distributed
steps = 30
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
t0 = time.perf_counter()
s0.record()
for i in range(steps): loop.resident_step(i)
s1.record()
t_enq = time.perf_counter() - t0
torch.cuda.synchronize()
tot = s0.elapsed_time(s1) / steps
by = {}
for n, a, b in marks:
    by.setdefault(n, []).append(a.elapsed_time(b))
print(rank, mode, f'step {tot:.3f} ms; host enqueue {t_enq / steps * 1e3:.3f} ms/step;', {n: (len(v) // steps, round(sum(v) / len(v), 3)) for n, v in by.items()}, flush=True)
dist.all_reduce = orig
dist.barrier(); dist.destroy_process_group()
