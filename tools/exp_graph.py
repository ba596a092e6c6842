"""Experiment: how much of the step is launch gaps / host overhead?  Captures one full SASRec step
(fixed dropout step) in a CUDA graph and times replays against the eager public-API loop."""
import sys, os, time
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
batch = {k: v.to(dev) for k, v in synthetic_batch(B, 50, N, seed=1).items()}
eng = m.engine
table = m.item_embedding.weight.data
opt = m.optimizer

def eager():
    b = dict(batch); b['neg_item'] = m._neg_sampling(b)
    opt.zero_grad(); loss = m.training_step(b); loss.backward(); opt.step()

def raw_step():
    neg = batch['neg_item'].view(B, 50)
    bufs = eng.prep(batch['seqlen'], batch['item_id'])
    eng.encode(bufs, table, m._flat, batch['in_item_id'], train=True)
    eng.score_bce(bufs, table, batch['item_id'], neg, want_grad=True)
    eng.reduce_loss(bufs)
    eng.encode_bwd(bufs, table, m._flat, batch['in_item_id'], m._flat_grad)
    eng.table_grad(bufs, batch['in_item_id'], batch['item_id'], neg, m._table_grad, m._flat_grad[:50 * D].view(50, D))
    E.adam_step(m._flat, m._flat_grad, opt.flat_groups[0].m, opt.flat_groups[0].v, 5, 1e-3)
    E.adam_step(table, m._table_grad, opt.flat_groups[1].m, opt.flat_groups[1].v, 5, 1e-3, zero_grad=True)

def timeit(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b_.record(); torch.cuda.synchronize()
    return a.elapsed_time(b_) / n

print('eager public API  ms/step', round(timeit(eager), 4))
print('raw engine calls  ms/step', round(timeit(raw_step), 4))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): raw_step()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=s):
    raw_step()
print('graph replay      ms/step', round(timeit(g.replay), 4))

# ---- attention backend comparison (raw engine calls + per-kernel profile) ----
import ctypes
lib = E._lib.lib()
for backend in ():
    lib.dr4sr_set_attn_backend(backend)
    print(f'attn backend {backend}: raw ms/step', round(timeit(raw_step), 4))
    lib.dr4sr_prof_enable(1)
    for _ in range(20): raw_step()
    torch.cuda.synchronize()
    lib.dr4sr_prof_enable(0)
    buf = ctypes.create_string_buffer(1 << 16)
    lib.dr4sr_prof_collect(buf, len(buf))
    for line in buf.value.decode().strip().splitlines():
        name, cnt, tot = line.split(',')
        if 'attn' in name: print('   ', name, cnt, round(float(tot) / 20 * 1000, 1), 'us/step')
pass
