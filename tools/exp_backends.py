"""Per-kernel timing of one SASRec step under a chosen (attn_backend, fused) pair."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dr4sr_b200.data.synthetic import synthetic_batch
from dr4sr_b200.model.sasrec import SASRec
from dr4sr_b200.utils.config import SyntheticCatalog, default_config
from dr4sr_b200 import engine as E
dev = 'cuda:0'
N, D, B = 100_000, 128, 1024
cfg = default_config('SASRec', model__embed_dim=D, train__device=dev, train__batch_size=B)
torch.manual_seed(0)
m = SASRec(cfg, [SyntheticCatalog(N)] * 3); m._init_model(); m.train()
lib = E._lib.lib()
batches = [{k: v.to(dev) for k, v in synthetic_batch(B, 50, N, seed=i).items()} for i in range(4)]
def step(i):
    b = dict(batches[i % 4]); b['neg_item'] = m._neg_sampling(b)
    m.optimizer.zero_grad(); loss = m.training_step(b); loss.backward(); m.optimizer.step()
def timeit(n=100):
    for i in range(10): step(i)
    torch.cuda.synchronize()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): step(i)
    b_.record(); torch.cuda.synchronize()
    return a.elapsed_time(b_) / n
for attn, fused in ((0, 1), (1, 1)):
    lib.dr4sr_set_attn_backend(attn); lib.dr4sr_set_fused_backend(fused)
    t = timeit()
    lib.dr4sr_prof_enable(1)
    for i in range(20): step(i)
    torch.cuda.synchronize(); lib.dr4sr_prof_enable(0)
    buf = ctypes.create_string_buffer(1 << 16); lib.dr4sr_prof_collect(buf, len(buf))
    rows = [l.split(',') for l in buf.value.decode().strip().splitlines()]
    print(f'attn={attn} fused={fused}: {t:.4f} ms/step;', ', '.join(f'{n} {float(tot) / 20 * 1000:.0f}us' for n, c, tot in rows[:7]))
