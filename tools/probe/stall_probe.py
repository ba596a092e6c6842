"""N-rank probe: which host call stalls in the first few hundred steps of a multi-GPU run?  Per-step host time by phase, GC log."""
import os, sys, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
import bench
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local); dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
mode = sys.argv[1] if len(sys.argv) > 1 else 'peer'
if len(sys.argv) > 2 and sys.argv[2] == 'nogc': gc.disable()
c = dict(bench.CFG2); c['adam_rows'] = c['num_items'] // world if mode == 'peer' else c['num_items']
model, layout = bench.build_model('SASRec', c, dev, rank, world, sharded=mode == 'peer')
if mode == 'peer': model.enable_peer_table(dist.group.WORLD)
else: model.enable_data_parallel(dist.group.WORLD)
model.train()
loop = bench.StepLoop(model, layout, c, dev, rank)
gclog = []
def cb(phase, info):
    if phase == 'start': cb.t = time.perf_counter()
    else: gclog.append((cb.step, info['generation'], (time.perf_counter() - cb.t) * 1e3))
cb.step = -1
gc.callbacks.append(cb)
m = loop.model
rows = []
mem0 = torch.cuda.memory_stats()['num_device_alloc'] if hasattr(torch.cuda, 'memory_stats') else 0
for i in range(400):
    cb.step = i
    t = [time.perf_counter()]
    batch = dict(loop.resident[i % loop.P]); batch['neg_item'] = m._neg_sampling(batch); t.append(time.perf_counter())
    m.optimizer.zero_grad(); t.append(time.perf_counter())
    loss = m.training_step(batch=batch); t.append(time.perf_counter())
    loss.backward(); t.append(time.perf_counter())
    m.optimizer.step(); t.append(time.perf_counter())
    rows.append((i, [(t[j + 1] - t[j]) * 1e3 for j in range(5)], torch.cuda.memory_stats()['num_device_alloc']))
torch.cuda.synchronize()
slow = sorted(rows, key=lambda r: -sum(r[1]))[:6]
print(rank, mode, 'slowest steps (neg, zero, fwd, bwd, adam ms; cudaMallocs so far):', [(i, [round(x, 1) for x in ph], n - mem0) for i, ph, n in sorted(slow)], flush=True)
print(rank, 'gc gen>=1 events (step, gen, ms):', [(s, g, round(ms, 1)) for s, g, ms in gclog if g >= 1 or ms > 2][:20], flush=True)
dist.barrier(); dist.destroy_process_group()
