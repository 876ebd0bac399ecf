"""Where the end-to-end step's time goes (C2 batch): eager resident step, + structure build per step, + loader.
    python scripts/prof_e2e.py"""
import itertools, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CFG, make_graphs
graphs = make_graphs(list(range(25)), 143107, 1_500_000)
import torch
from meta_gcn_b200 import dist as mdist, functional as F
from meta_gcn_b200.data import DeviceLoader, GraphBatch
from meta_gcn_b200.gcn_meta.models import GCNModel
from meta_gcn_b200.graph import clear_structure_cache
dev = torch.device("cuda")
torch.cuda.set_stream(torch.cuda.Stream(dev))
host = GraphBatch.from_data_list(graphs).with_int32_indices().pin_memory()
torch.manual_seed(0)
model = GCNModel(**CFG).to(dev)
reducer = mdist.FlatGradientReducer(model.parameters())
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-4, fused=True)

def step(b, sync=True):
    reducer.zero()
    out = model(b.x[:, 0].view(-1, 1), b.edge_index, deg_K=b.x[:, 1])
    loss = F.cross_entropy(out, b.y.long(), "sum")
    loss.backward()
    m, _ = reducer.reduce_mean(loss, b.num_nodes)
    opt.step()
    return float(m.item()) if sync else m

def wall(fn, k=10):
    fn(); fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k * 1e3

b = host.to(dev)
b.x = b.x.contiguous()
print(f"eager resident step, loss read every step : {wall(lambda: step(b)):.2f} ms")
print(f"eager resident step, no host read          : {wall(lambda: step(b, False)):.2f} ms")
def with_build():
    clear_structure_cache()
    b.structure()
    return step(b)
print(f"+ structure build every step               : {wall(with_build):.2f} ms")
t0 = time.perf_counter(); 
for _ in range(10):
    reducer.zero(); out = model(b.x[:, 0].view(-1, 1), b.edge_index, deg_K=b.x[:, 1])
t_host = (time.perf_counter() - t0) / 10 * 1e3
torch.cuda.synchronize()
print(f"host time to ENQUEUE a forward             : {t_host:.2f} ms")
loader = DeviceLoader((), dev, fields=("x", "edge_index", "y"))
def e2e(k):
    loader.batches = itertools.repeat(host, k)
    for bb in loader:
        step(bb)
e2e(2); torch.cuda.synchronize()
t0 = time.perf_counter(); e2e(20); torch.cuda.synchronize()
print(f"loader e2e (20 steps)                      : {(time.perf_counter() - t0) / 20 * 1e3:.2f} ms")
