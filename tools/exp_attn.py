"""Experiment driver for ncu: a few raw SASRec steps with a chosen attention backend."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dr4sr_b200.data.synthetic import synthetic_batch
from dr4sr_b200.model.sasrec import SASRec
from dr4sr_b200.utils.config import SyntheticCatalog, default_config
from dr4sr_b200 import engine as E
backend = int(sys.argv[1]) if len(sys.argv) > 1 else 0
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = 'cuda:0'
N, D, B = 100_000, 128, 1024
cfg = default_config('SASRec', model__embed_dim=D, train__device=dev, train__batch_size=B)
torch.manual_seed(0)
m = SASRec(cfg, [SyntheticCatalog(N)] * 3); m._init_model(); m.train()
E._lib.lib().dr4sr_set_attn_backend(backend)
batch = {k: v.to(dev) for k, v in synthetic_batch(B, 50, N, seed=1).items()}
for _ in range(steps):
    b = dict(batch); b['neg_item'] = m._neg_sampling(b)
    m.optimizer.zero_grad(); loss = m.training_step(b); loss.backward(); m.optimizer.step()
torch.cuda.synchronize()
print('ok', float(loss))
