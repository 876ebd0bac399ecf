"""Secondary measurements for BASELINE.json configs[0], [2], [3], [4] (the bench.py line is configs[1]).
CUDA events, synthetic data of the named shapes; one JSON object per line.
    python scripts/bench_configs.py [--skip-c4] > gpurun_out/configs.jsonl
The "torch path" columns time the reference's own op sequence on the same GPU with stock PyTorch ops
(index_select -> mul -> index_add_, gcn_base_models.py:223-237) — what the reference would run on CUDA."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--skip-c4", action="store_true")
ap.add_argument("--only-c4", action="store_true")
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()

import numpy as np  # noqa: E402
import torch  # noqa: E402

from meta_gcn_b200 import data as D  # noqa: E402
from meta_gcn_b200 import functional as F_mgcn  # noqa: E402
from meta_gcn_b200 import kernel as K  # noqa: E402
from meta_gcn_b200 import ops  # noqa: E402
from meta_gcn_b200.gcn_meta.models import GCNModel  # noqa: E402
from meta_gcn_b200.graph import GraphStructure, clear_structure_cache  # noqa: E402

dev = torch.device("cuda")


def timeit(fn, reps=args.reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def emit(**kw):
    print(json.dumps(kw), flush=True)


if args.only_c4:
    C1 = C3 = C5 = False
else:
    C1 = C3 = C5 = True

# ---- C1: one botnet graph, 12-layer residual GCN, fwd + loss + bwd (latency-bound: 44 MB/layer < L2) ----
if C1:
    g = D.synth_botnet_graph(seed=0)
    b = D.GraphBatch.from_data_list([g]).to(dev)
    cfg = dict(in_channels=1, enc_sizes=[32] * 12, num_classes=2, residual_hop=1, dropout=0.0, final_type="proj",
               deg_norm="sm", aggr="add", bias=False)
    torch.manual_seed(0)
    model = GCNModel(**cfg).to(dev)
    x0, deg, y = b.x[:, 0].reshape(-1, 1).contiguous(), b.x[:, 1].contiguous(), b.y.long()


    def c1_step():
        model.zero_grad(set_to_none=True)
        out = model(x0, b.edge_index, deg_K=deg)
        F_mgcn.cross_entropy(out, y, "mean").backward()


    ms = timeit(c1_step, reps=10, warm=3)
    ops_l0 = ops._lib.launch_count()
    c1_step()
    launches = ops._lib.launch_count() - ops_l0
    emit(config="C1", what="12-layer residual GCN h=32, one botnet graph, fwd+loss+bwd (structures cached)", ms=ms,
         graphs_per_s=1e3 / ms, gedges_per_s=b.num_edges * 24 / ms / 1e6, launches=launches,
         n=b.num_nodes, e=b.num_edges)

    # the same step replayed as one CUDA graph (meta_gcn_b200/graphed.py): C1 is launch-bound
    from meta_gcn_b200.graphed import GraphedCall  # noqa: E402
    for p_ in model.parameters():
        p_.grad = torch.zeros_like(p_)


    def c1_fwd_loss_bwd():
        for p_ in model.parameters():
            p_.grad.zero_()
        loss = F_mgcn.cross_entropy(model(x0, b.edge_index, deg_K=deg), y, "mean")
        loss.backward()
        return loss


    graphed = GraphedCall(c1_fwd_loss_bwd)
    ms_g = timeit(graphed, reps=20, warm=3)
    emit(config="C1", what="the same step replayed as one CUDA graph", ms=ms_g, graphs_per_s=1e3 / ms_g,
         gedges_per_s=b.num_edges * 24 / ms_g / 1e6, n=b.num_nodes, e=b.num_edges)

# ---- C3: TU-shaped batches of 128 small graphs, 3 layers, hidden 64 ----
if C3:
    tb = D.synth_tu_batch(seed=0, num_graphs=128).to(dev)
    meta = D.dataset_meta(3, 2)
    for name, cls in (("GCN", K.GCN), ("GIN0", K.GIN0), ("GraphSAGE", K.GraphSAGE)):
        torch.manual_seed(0)
        net = cls(meta, 3, 64).to(dev).train()
        yb = tb.y.view(-1).long()

        def c3_step():
            clear_structure_cache()          # a new mini-batch every step: structure build included
            net.zero_grad(set_to_none=True)
            out = net(tb)
            torch.nn.functional.nll_loss(out, yb).backward()

        ms = timeit(c3_step, reps=20, warm=5)
        emit(config="C3", what=f"kernel/{name} 3 layers hidden 64, batch of 128 TU-shaped graphs, fwd+bwd incl. structure build",
             ms=ms, graphs_per_s=128e3 / ms, n=tb.num_nodes, e=tb.num_edges)

# ---- C5: aggregation sweep vs the reference's op sequence with stock torch CUDA ops ----
if C5:
    N5 = 1_000_000
    for avg_deg in (2, 8, 32, 128):
        ei = torch.from_numpy(D.synth_powerlaw_graph(5, N5, N5 * avg_deg)).to(dev)
        gs = GraphStructure(ei, N5)
        gs.fwd
        norm = torch.rand(ei.size(1), device=dev)
        for H in (16, 32, 64, 128, 256, 512):
            x = torch.randn(N5, H, device=dev)
            ours = timeit(lambda: ops.spmm_impl(gs.fwd, x), reps=3, warm=1)
            plain = timeit(lambda: ops.aggregate_prescaled_impl(gs.fwd, x), reps=3, warm=1) if H in (16, 32, 64, 128) else None

            def torch_path():
                xj = x.index_select(0, ei[0]) * norm.view(-1, 1)
                return torch.zeros(N5, H, device=dev).index_add_(0, ei[1], xj)

            ref = timeit(torch_path, reps=2, warm=1) if ei.size(1) * H * 4 < 40e9 else None
            emit(config="C5", avg_degree=avg_deg, H=H, n=N5, e=int(ei.size(1)), spmm_ms=ours, plain_ms=plain,
                 torch_index_add_ms=ref, speedup_vs_torch=(ref / min(ours, plain or ours)) if ref else None,
                 gather_gbs=ei.size(1) * H * 4 / min(ours, plain or ours) / 1e6)
            del x
        del gs, ei, norm
        clear_structure_cache()
        torch.cuda.empty_cache()

# ---- C4: ogbn-products-shaped graph, GraphSAGE hidden 256 (tcgen05 wide transform) ----
if not args.skip_c4:
    N4, E4 = 2_449_029, 61_859_140
    ei = torch.from_numpy(D.synth_powerlaw_graph(4, N4, E4, alpha=0.6)).to(dev)
    x = torch.randn(N4, 100, device=dev)
    yb = torch.randint(0, 47, (1,), device=dev)
    batch = D.GraphBatch(x, ei, yb, torch.zeros(N4, dtype=torch.long, device=dev), [0, N4])
    torch.manual_seed(0)
    net = K.GraphSAGE(D.dataset_meta(100, 47), 3, 256).to(dev).train()

    def c4_fwd():
        with torch.no_grad():
            return net(batch)

    def c4_step():
        net.zero_grad(set_to_none=True)
        torch.nn.functional.nll_loss(net(batch), yb).backward()

    ms_f = timeit(c4_fwd, reps=3, warm=1)
    ms_s = timeit(c4_step, reps=3, warm=1)
    emit(config="C4", what="kernel/GraphSAGE 3 layers hidden 256 on a products-shaped graph (structures cached)",
         fwd_ms=ms_f, fwd_bwd_ms=ms_s, n=N4, e=E4, gedges_per_s_fwd=E4 * 3 / ms_f / 1e6,
         peak_mem_gb=torch.cuda.max_memory_allocated() / 1e9)
