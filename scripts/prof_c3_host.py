"""Host-side profile of one C3 step (kernel/gcn.py, 128 TU-shaped graphs): where the 1.6 ms go.
    python scripts/prof_c3_host.py"""
import os, sys, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from meta_gcn_b200 import _lib, data as D, kernel as K
from meta_gcn_b200.graph import clear_structure_cache
dev = torch.device("cuda")
tb = D.synth_tu_batch(seed=0, num_graphs=128).to(dev)
meta = D.dataset_meta(3, 2)
torch.manual_seed(0)
net = K.GCN(meta, 3, 64).to(dev).train()
yb = tb.y.view(-1).long()
def step():
    clear_structure_cache()
    net.zero_grad(set_to_none=True)
    torch.nn.functional.nll_loss(net(tb), yb).backward()
for _ in range(5):
    step()
torch.cuda.synchronize()
c0 = _lib.launch_count()
step()
torch.cuda.synchronize()
print("libmgcn launches per step:", _lib.launch_count() - c0)
import time
t0 = time.perf_counter()
for _ in range(50):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0)/50:.3f} ms per step; with final sync {1e3*(t2-t0)/50:.3f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
