"""Structure build (edge_index -> row structures + work order) and H2D timing at the C2 shape: the radix-sort build,
the sort-free build of an already ordered list, the layout check with its host read.
    python scripts/prof_build.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_graphs
graphs = make_graphs(list(range(25)), 143107, 1_500_000)
import torch
from meta_gcn_b200 import ops
from meta_gcn_b200.data import GraphBatch
from meta_gcn_b200.graph import GraphStructure
dev = torch.device("cuda")
host = GraphBatch.from_data_list(graphs).pin_memory()
host32 = host.with_int32_indices().pin_memory()
b = host.to(dev)
b32 = host32.to(dev)
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

def wall(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3

print(f"N={n} E={e}")
for name, ei in (("int64", b.edge_index), ("int32", b32.edge_index)):
    print(f"[{name}] csr_build by target, radix sort : {t(lambda: ops.csr_build_impl(ei, n, 1, 0, 64)):.3f} ms")
    print(f"[{name}] csr_build by target, presorted  : {t(lambda: ops.csr_build_impl(ei, n, 1, 0, 64, layout=2)):.3f} ms")
    print(f"[{name}] csr_build by source, presorted  : {t(lambda: ops.csr_build_impl(ei, n, 0, 0, 64, layout=2)):.3f} ms")
    print(f"[{name}] edge_layout (+ host read)       : {wall(lambda: ops.edge_layout_impl(ei, n)):.3f} ms wall")
    print(f"[{name}] GraphStructure.fwd end to end    : {wall(lambda: GraphStructure(ei, n).fwd):.3f} ms wall")
    print(f"[{name}] GraphStructure.fwd_plain         : {wall(lambda: GraphStructure(ei, n).fwd_plain):.3f} ms wall")
    def both():
        g_ = GraphStructure(ei, n); g_.fwd_plain; g_.bwd_plain
    print(f"[{name}] fwd_plain + bwd_plain            : {wall(both):.3f} ms wall")
print(f"H2D pinned batch int64 ({(host.edge_index.numel()*8 + host.x.numel()*4)/1e6:.0f} MB): {t(lambda: host.to(dev, non_blocking=True)):.3f} ms")
print(f"H2D pinned batch int32 ({(host32.edge_index.numel()*4 + host32.x.numel()*4)/1e6:.0f} MB): {t(lambda: host32.to(dev, non_blocking=True)):.3f} ms")
