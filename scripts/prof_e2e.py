"""Where the e2e step goes beyond the resident step: variants of bench.py's e2e loop at the C2 shape.
    python scripts/prof_e2e.py"""
import itertools
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CFG, make_graphs  # noqa: E402

graphs = make_graphs(list(range(25)), 143107, 1_500_000)
import torch  # noqa: E402
from meta_gcn_b200 import functional as F  # noqa: E402
from meta_gcn_b200.data import DeviceLoader, GraphBatch  # noqa: E402
from meta_gcn_b200.gcn_meta.models import GCNModel  # noqa: E402
from meta_gcn_b200.graph import clear_structure_cache, structure_of  # noqa: E402

dev = torch.device("cuda")
host = GraphBatch.from_data_list(graphs).pin_memory()
torch.manual_seed(0)
model = GCNModel(**CFG).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-4, fused=True)


def step(b):
    opt.zero_grad(set_to_none=False)
    out = model(b.x[:, 0].view(-1, 1), b.edge_index, deg_K=b.x[:, 1])
    loss = F.cross_entropy(out, b.y.long(), "sum")
    loss.backward()
    opt.step()
    return loss


def timed(name, fn, k=12):
    fn(2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn(k)
    torch.cuda.synchronize()
    print(f"{name:58s} {(time.perf_counter() - t0) / k * 1e3:8.2f} ms/step", flush=True)


resident = host.to(dev)
loader = DeviceLoader((), dev)


def v_resident(k):
    for _ in range(k):
        step(resident)


def v_resident_sync(k):
    for _ in range(k):
        float(step(resident).item())


def v_resident_rebuild(k):
    for _ in range(k):
        clear_structure_cache()
        float(step(resident).item())


def v_loader_nobuild(k):
    loader.batches = itertools.repeat(host, k)
    for b in loader:
        float(step(b).item())


def v_full(k):
    loader.batches = itertools.repeat(host, k)
    for b in loader:
        clear_structure_cache()
        float(step(b).item())


timed("resident batch, structures cached, no per-step sync", v_resident)
timed("  + loss.item() every step", v_resident_sync)
timed("  + structure rebuilt every step", v_resident_rebuild)
timed("DeviceLoader H2D every step, structures cached (slot reuse)", v_loader_nobuild)
timed("DeviceLoader + rebuild (= bench e2e)", v_full)
gs = structure_of(resident.edge_index, resident.num_nodes)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    clear_structure_cache()
    g2 = structure_of(resident.edge_index, resident.num_nodes)
    g2.fwd
    g2.symmetric
torch.cuda.synchronize()
print(f"structure build (fwd) + symmetry check alone                  {(time.perf_counter() - t0) / 5 * 1e3:8.2f} ms")
