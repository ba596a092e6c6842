"""Per-phase cycle averages of the tcgen05 GRU recurrence (needs the trace build: make variant V=trace X=-DDR4SR_TRACE)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ['DR4SR_LIB_PATH'] = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'dr4sr_b200/csrc/libdr4sr_trace.so')
import torch
import bench
dev = torch.device('cuda', 0)
c = dict(bench.CFG2)
model, layout = bench.build_model('GRU4Rec', c, dev, 0, 1, False)
model.train()
loop = bench.StepLoop(model, layout, c, dev, 0)
for i in range(2):
    loop.resident_step(i)
torch.cuda.synchronize()
