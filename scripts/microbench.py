"""Kernel micro-benchmarks at the C2 shape (CUDA events, inputs larger than L2).
    python scripts/microbench.py [--graphs 25]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--graphs", type=int, default=25)
ap.add_argument("--hub", type=int, default=256)
args = ap.parse_args()

from bench import make_graphs, b_agg  # noqa: E402

graphs = make_graphs(list(range(args.graphs)), 143107, 1_500_000)
import torch  # noqa: E402
from meta_gcn_b200 import ops  # noqa: E402
from meta_gcn_b200.data import GraphBatch  # noqa: E402
from meta_gcn_b200.graph import GraphStructure  # noqa: E402

dev = torch.device("cuda")
b = GraphBatch.from_data_list(graphs).to(dev)
n, e = b.num_nodes, b.num_edges
H = 32


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps


t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
gs = GraphStructure(b.edge_index, n, hub_threshold=args.hub)
gs.fwd, gs.bwd
t1.record(); torch.cuda.synchronize()
print(f"structure build (fwd+bwd): {t0.elapsed_time(t1):.2f} ms  hubs={int(gs.fwd.hub_count)} segs={int(gs.fwd.seg_count)}")
dis = ops.gcn_norm_impl(b.x[:, 1].contiguous(), 0)
x = torch.randn(n, H, device=dev)
w = torch.randn(H, H, device=dev)
bytes_agg = b_agg(n, e, H)
print(f"hub threshold {args.hub}")
for name, fn in [
    ("spmm exact (sm weights)", lambda: ops.spmm_impl(gs.fwd, x, nbr_scale=dis, row_scale=dis, act=1)),
    ("spmm plain (no weights)", lambda: ops.spmm_impl(gs.fwd, x)),
    ("aggregate_prescaled fwd", lambda: ops.aggregate_prescaled_impl(gs.fwd, x, dis, 0, None, None, 1)),
    ("aggregate_prescaled bwd", lambda: ops.aggregate_prescaled_impl(gs.bwd, x, dis, 0, None, None, 0)),
]:
    ms = timeit(fn)
    print(f"{name:32s} {ms:8.3f} ms  {bytes_agg / ms / 1e6:8.1f} GB/s algorithmic")
g = torch.randn(n, H, device=dev)
for name, fn, byts in [
    ("linear 32x32", lambda: ops.linear_impl(x, w, False), 8 * n * H),
    ("linear 32x32 +add+relu", lambda: ops.linear_impl(x, w, True, None, g, 1), 12 * n * H),
    ("wgrad 32x32", lambda: ops.linear_wgrad_impl(x, g, False, True), 8 * n * H),
    ("relu_backward", lambda: ops.relu_backward_impl(g, x), 12 * n * H),
    ("torch copy (reference)", lambda: g.copy_(x), 8 * n * H),
]:
    ms = timeit(fn)
    print(f"{name:32s} {ms:8.3f} ms  {byts / ms / 1e6:8.1f} GB/s")
