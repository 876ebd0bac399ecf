"""Structure build (edge_index -> CSR/CSC + work order) and H2D timing at the C2 shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_graphs
graphs = make_graphs(list(range(25)), 143107, 1_500_000)
import torch
from meta_gcn_b200 import ops
from meta_gcn_b200.data import GraphBatch
dev = torch.device("cuda")
host = GraphBatch.from_data_list(graphs).pin_memory()
b = host.to(dev)
n, e = b.num_nodes, b.num_edges

def t(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

print(f"N={n} E={e}")
print(f"csr_build by target: {t(lambda: ops.csr_build_impl(b.edge_index, n, 1, 0, 64)):.3f} ms")
print(f"csr_build by source: {t(lambda: ops.csr_build_impl(b.edge_index, n, 0, 0, 64)):.3f} ms")
print(f"H2D pinned batch ({(host.edge_index.numel()*8 + host.x.numel()*4)/1e6:.0f} MB): {t(lambda: host.to(dev, non_blocking=True)):.3f} ms")
