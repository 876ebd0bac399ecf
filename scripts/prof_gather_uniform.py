"""Plain hidden-32 aggregation on a batch with the botnet shape but UNIFORM endpoints inside every graph (no hubs),
next to the power-law batch of bench.py: how much of the gather time is the hub-heavy index stream?
    python scripts/prof_gather_uniform.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meta_gcn_b200 import ops  # noqa: E402
from meta_gcn_b200.graph import GraphStructure  # noqa: E402

dev = torch.device("cuda")
G, n, e = 25, 143107, 1_500_000
rng = np.random.default_rng(0)
parts = []
for g in range(G):
    src = rng.integers(0, n, e) + g * n
    dst = rng.integers(0, n, e) + g * n
    parts.append(np.stack([src, dst]))
ei = torch.from_numpy(np.concatenate(parts, axis=1)).to(dev)
N = G * n
x = torch.randn(N, 32, device=dev)
gs = GraphStructure(ei, N)
gs.fwd


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps


ms = timeit(lambda: ops.aggregate_prescaled_impl(gs.fwd, x, None, 0, None, None, 0))
print(f"uniform endpoints: N={N} E={ei.size(1)} hubs={int(gs.fwd.hub_count)}  {ms:.3f} ms  "
      f"{ei.size(1) * 128 / ms / 1e9:.2f} TB/s of gathered rows")
