"""N-rank probe: device time per step of the resident loop as a function of how far the host may run ahead of the GPU
(unbounded / at most K steps), with and without the nvidia-smi sampler bench.py runs.  Usage: runahead_probe.py {peer|replicated}"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
import bench
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local); dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
mode = sys.argv[1] if len(sys.argv) > 1 else 'peer'
c = dict(bench.CFG2); c['adam_rows'] = c['num_items'] // world if mode == 'peer' else c['num_items']
model, layout = bench.build_model('SASRec', c, dev, rank, world, sharded=mode == 'peer')
if mode == 'peer': model.enable_peer_table(dist.group.WORLD)
else: model.enable_data_parallel(dist.group.WORLD)
model.train()
loop = bench.StepLoop(model, layout, c, dev, rank)
for i in range(30): loop.resident_step(i)
torch.cuda.synchronize(); dist.barrier()

def run(steps, depth, sampler):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps // 10 + 1)]
    ring = [torch.cuda.Event() for _ in range(max(depth, 1))]
    ctx = bench.ClockSampler(local) if sampler else None
    if ctx: ctx.__enter__(); time.sleep(0.3)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter(); host = 0.0
    evs[0].record()
    for i in range(steps):
        if depth and i >= depth:
            ring[i % depth].synchronize()
        h0 = time.perf_counter()
        loop.resident_step(i)
        host += time.perf_counter() - h0
        if depth: ring[i % depth].record()
        if (i + 1) % 10 == 0: evs[(i + 1) // 10].record()
    torch.cuda.synchronize()
    if ctx: ctx.__exit__()
    seg = [evs[j].elapsed_time(evs[j + 1]) / 10 for j in range(steps // 10)]
    t = torch.tensor([evs[0].elapsed_time(evs[-1]) / steps, host / steps * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f'{mode} N={world} depth={depth or "inf"} sampler={int(sampler)}: {float(t[0]):.3f} ms/step (max over ranks), host enqueue {float(t[1]):.3f} ms/step; '
              f'per-10-step ms: {" ".join(f"{x:.2f}" for x in seg)}', flush=True)

for depth, sampler in ((0, False), (2, False), (4, False), (0, True), (2, True), (0, False)):
    run(100, depth, sampler)
dist.barrier(); dist.destroy_process_group()
