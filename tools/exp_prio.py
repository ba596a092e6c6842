"""Does running the step on a high-priority stream (side stream stays at default priority) shorten the critical path?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dr4sr_b200.data.synthetic import synthetic_batch
from dr4sr_b200.model.sasrec import SASRec
from dr4sr_b200.utils.config import SyntheticCatalog, default_config
dev = 'cuda:0'
N, D, B = 100_000, 128, 1024
cfg = default_config('SASRec', model__embed_dim=D, train__device=dev, train__batch_size=B)
torch.manual_seed(0)
m = SASRec(cfg, [SyntheticCatalog(N)] * 3); m._init_model(); m.train()
batches = [{k: v.to(dev) for k, v in synthetic_batch(B, 50, N, seed=i).items()} for i in range(4)]
def step(i):
    b = dict(batches[i % 4]); b['neg_item'] = m._neg_sampling(b)
    m.optimizer.zero_grad(); loss = m.training_step(b); loss.backward(); m.optimizer.step()
def timeit(n=200):
    for i in range(20): step(i)
    torch.cuda.synchronize()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): step(i)
    b_.record(); torch.cuda.synchronize()
    return a.elapsed_time(b_) / n
print('default stream      ms/step', round(timeit(), 4))
for pr in (-1, -5):
    s = torch.cuda.Stream(priority=pr)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        print(f'stream priority {pr}  ms/step', round(timeit(), 4))
