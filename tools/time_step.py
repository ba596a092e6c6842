"""Step / forward timing of the cfg-2 SASRec step for the library selected by DR4SR_LIB_PATH, with the per-kernel event profile.
usage: python tools/time_step.py [attn_backend] [fused_backend]"""
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
if len(sys.argv) > 1: lib.dr4sr_set_attn_backend(int(sys.argv[1]))
if len(sys.argv) > 2: lib.dr4sr_set_fused_backend(int(sys.argv[2]))
batches = [{k: v.to(dev) for k, v in synthetic_batch(B, 50, N, seed=i).items()} for i in range(4)]
def step(i):
    b = dict(batches[i % 4]); b['neg_item'] = m._neg_sampling(b)
    m.optimizer.zero_grad(); loss = m.training_step(b); loss.backward(); m.optimizer.step()
def timeit(fn, n=100):
    for i in range(10): fn(i)
    torch.cuda.synchronize()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): fn(i)
    b_.record(); torch.cuda.synchronize()
    return a.elapsed_time(b_) / n
t = timeit(step)
tf = timeit(lambda i: m.forward(batches[i % 4]))
lib.dr4sr_prof_enable(1)
for i in range(20): step(i)
torch.cuda.synchronize(); lib.dr4sr_prof_enable(0)
buf = ctypes.create_string_buffer(1 << 16); lib.dr4sr_prof_collect(buf, len(buf))
rows = [l.split(',') for l in buf.value.decode().strip().splitlines()]
print(f'{os.environ.get("DR4SR_LIB_PATH", "default")}: step {t:.4f} ms; forward-only {tf:.4f} ms;', ', '.join(f'{n} {float(tot) / 20 * 1000:.0f}us' for n, c, tot in rows[:12]))
