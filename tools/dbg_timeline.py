"""Timeline of the fused forward kernel (library built with -DDR4SR_TRACE): per-phase cycle counts of the first CTAs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dr4sr_b200 import _lib
from dr4sr_b200.data.synthetic import synthetic_batch
from dr4sr_b200.model.sasrec import SASRec
from dr4sr_b200.utils.config import SyntheticCatalog, default_config
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = 'cuda:0'
N, D = 100000, 128
cfg = default_config('SASRec', model__embed_dim=D, model__dropout_rate=0.5, train__device=dev, train__batch_size=B)
torch.manual_seed(0)
m = SASRec(cfg, [SyntheticCatalog(N)] * 3); m._init_model(); m.train()
lib = _lib.lib()
batch = {k: v.to(dev) for k, v in synthetic_batch(B, 50, N, seed=1).items()}
for _ in range(3):
    m.forward(batch)
torch.cuda.synchronize()
tr = torch.full((8 * 256,), -1, dtype=torch.int32).pin_memory()
_lib.check(lib.dr4sr_debug_trace(tr.data_ptr()), 'trace')
m.forward(batch)
torch.cuda.synchronize()
lib.dr4sr_debug_trace(None)
import collections
for cta in (0, 5):
    t = tr.view(8, 128, 2)[cta]
    if len(sys.argv) > 2:
        pv = 0
        for code, cyc in t.tolist()[:int(sys.argv[2])]:
            print(f'  {code:5d} t={cyc:8d} (+{cyc - pv})'); pv = cyc
        continue
    prev, prevc = 0, 0
    agg = collections.OrderedDict()
    last = 0
    for code, cyc in t.tolist():
        if code < 0: break
        key = (prevc % 100 if prevc < 1000 else prevc, code % 100 if code < 1000 else code)
        a = agg.setdefault(key, [0, 0]); a[0] += 1; a[1] += cyc - prev
        prev, prevc, last = cyc, code, cyc
    print('CTA', cta, 'B', B, 'total cycles', last)
    for k, (n, c) in agg.items():
        print(f'   {k[0]:5d} -> {k[1]:5d}: n={n:2d} total={c:7d} avg={c // n}')
