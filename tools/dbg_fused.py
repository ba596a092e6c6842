"""Debug harness for the persistent fused kernels: runs ONE small forward (and optionally a training step) with a
host-mapped progress trace installed and a hard deadline; on a stall it prints where every CTA stopped and exits."""
import sys, os, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dr4sr_b200 import _lib
from dr4sr_b200.data.synthetic import synthetic_batch
from dr4sr_b200.model.sasrec import SASRec
from dr4sr_b200.utils.config import SyntheticCatalog, default_config

B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
mode = sys.argv[2] if len(sys.argv) > 2 else 'fwd'
dev = 'cuda:0'
N, D = 3000, 128
cfg = default_config('SASRec', model__embed_dim=D, model__dropout_rate=0.0, train__device=dev, train__batch_size=B)
torch.manual_seed(0)
m = SASRec(cfg, [SyntheticCatalog(N)] * 3); m._init_model(); m.train()
lib = _lib.lib()
tr = torch.full((4096,), -1, dtype=torch.int32).pin_memory()
_lib.check(lib.dr4sr_debug_trace(tr.data_ptr()), 'trace')
batch = {k: v.to(dev) for k, v in synthetic_batch(B, 50, N, seed=1).items()}
torch.cuda.synchronize()
done = torch.cuda.Event()
if mode == 'fwd':
    q = m.forward(batch)
else:
    m.optimizer.zero_grad(); loss = m.training_step(batch); loss.backward()
done.record()
t0 = time.time()
while not done.query():
    if time.time() - t0 > 8:
        t = tr.view(-1, 4)
        live = t[t[:, 0] >= 0]
        print('STALL: phase histogram', collections.Counter(live[:, 0].tolist()))
        print('first rows', live[:8].tolist())
        sys.stdout.flush()
        os._exit(3)
    time.sleep(0.05)
try:
    torch.cuda.synchronize()
    t = tr.view(-1, 4)
    print('done; phases', collections.Counter(t[t[:, 0] >= 0][:, 0].tolist()))
    if mode == 'fwd':
        lib.dr4sr_set_fused_backend(0)
        q0 = m.forward(batch)
        print('max abs diff vs per-op', float((q - q0).abs().max()), 'scale', float(q0.abs().max()))
except Exception as e:
    t = tr.view(-1, 4)
    live = t[t[:, 0] >= 0]
    print('ERROR', str(e)[:200])
    print('phase histogram', collections.Counter(live[:, 0].tolist()))
    print('trapped rows', live[live[:, 1] >= 0][:8].tolist())
