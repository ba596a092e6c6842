"""Host-side cost of the training hooks: enqueue time per hook (no sync inside), and the e2e step with a loss read."""
import sys, os, time
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
batch = {k: v.to(dev) for k, v in synthetic_batch(B, 50, N, seed=1).items()}
T = {k: 0.0 for k in ('neg', 'zero', 'fwd', 'bwd', 'opt', 'sync')}
n = 300
for it in range(n + 20):
    if it == 20:
        T = {k: 0.0 for k in T}
    t0 = time.perf_counter(); b = dict(batch); b['neg_item'] = m._neg_sampling(b)
    t1 = time.perf_counter(); m.optimizer.zero_grad()
    t2 = time.perf_counter(); loss = m.training_step(b)
    t3 = time.perf_counter(); loss.backward()
    t4 = time.perf_counter(); m.optimizer.step()
    t5 = time.perf_counter(); x = float(loss)
    t6 = time.perf_counter()
    for k, d in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5)):
        T[k] += d
print({k: round(v / n * 1e6, 1) for k, v in T.items()}, 'us per step; total', round(sum(T.values()) / n * 1e6, 1))
