"""Timeline of the fused backward FFN-block kernel (library built with -DDR4SR_TRACE)."""
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
lib = _lib.lib(); lib.dr4sr_set_fused_backend(2)
batch = {k: v.to(dev) for k, v in synthetic_batch(B, 50, N, seed=1).items()}
def step():
    m.optimizer.zero_grad(); loss = m.training_step(batch); loss.backward(); m.optimizer.step()
for _ in range(3): step()
torch.cuda.synchronize()
tr = torch.full((8192,), -1, dtype=torch.int32).pin_memory()
_lib.check(lib.dr4sr_debug_trace(tr.data_ptr()), 'trace')
step()
torch.cuda.synchronize()
lib.dr4sr_debug_trace(None)
for cta in (0, 5):
    t = tr[:4096].view(16, 128, 2)[8 + cta]
    pv = 0
    print('CTA', cta)
    for code, cyc in t.tolist()[:60]:
        if code < 0: break
        print(f'  {code:5d} t={cyc:8d} (+{cyc - pv})'); pv = cyc
