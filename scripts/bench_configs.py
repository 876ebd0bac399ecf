"""Secondary measurements for BASELINE.json configs[0], [2], [3], [4] (the default bench.py line is configs[1]), each with
the reference's CPU path (oracle port, torch CPU fp32, all host threads) timed beside it on a bounded sample.
CUDA events, synthetic data of the named shapes.

    python scripts/bench_configs.py [--only 0,2] [--no-cpu]      one JSON object per line
    python bench.py --config K                                   the same through bench.py (one JSON line)

The "torch_index_add" columns of config 4 time the reference's own op sequence on the same GPU with stock PyTorch CUDA
ops (index_select -> mul -> index_add_, gcn_base_models.py:223-237) — what the reference would run on CUDA."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BOTNET = dict(in_channels=1, enc_sizes=[32] * 12, num_classes=2, residual_hop=1, dropout=0.0, final_type="proj",
              deg_norm="sm", aggr="add", bias=False)


def _timeit(fn, reps=5, warm=2):
    import torch
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


def _cpu_time(fn, reps=2, warm=1):
    for _ in range(warm):
        fn()
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


class _Data:
    pass


def config0(with_cpu=True):
    """configs[0]: one botnet graph, 12-layer residual GCN, fwd + loss + bwd (latency-bound: 44 MB per layer < L2)"""
    import torch
    from meta_gcn_b200 import _lib, data as D, functional as F
    from meta_gcn_b200.gcn_meta.models import GCNModel
    from meta_gcn_b200.graphed import GraphedCall
    dev = torch.device("cuda")
    torch.cuda.set_stream(torch.cuda.Stream(dev))
    g = D.synth_botnet_graph(seed=0)
    b = D.GraphBatch.from_data_list([g]).to(dev)
    torch.manual_seed(0)
    model = GCNModel(**BOTNET).to(dev)
    x0, deg, y = b.x[:, 0].reshape(-1, 1).contiguous(), b.x[:, 1].contiguous(), b.y.long()
    def step():
        # PyTorch's whole-network capture recipe (and the default of optimizer.zero_grad since 2.0): gradients are
        # dropped, not zeroed, so backward ASSIGNS them — no fill and no accumulation kernel per parameter (78 tiny
        # launches, 0.2 ms of a 1.75 ms step: profiles/r2_launches_c1.csv); under capture they land in the graph's pool
        model.zero_grad(set_to_none=True)
        loss = F.cross_entropy(model(x0, b.edge_index, deg_K=deg), y, "mean")
        loss.backward()
        return loss

    ms = _timeit(step, reps=10, warm=3)
    c0 = _lib.launch_count()
    step()
    launches = _lib.launch_count() - c0
    graphed = GraphedCall(step)
    ms_g = _timeit(graphed, reps=20, warm=3)
    n, e = b.num_nodes, b.num_edges
    bytes_step = 12 * (2 * (4 * e + 8 * n + 8 * n * 32) + 4 * n * 32)
    out = {"config": 0, "what": "configs[0]: 12-layer residual GCN h=32, ONE botnet graph, fwd+loss+bwd (structures cached)",
           "n": n, "e": e, "ms_eager": ms, "ms_cuda_graph": ms_g, "launches": launches,
           "graphs_per_s": 1e3 / ms_g, "gedges_per_s": e * 24 / ms_g / 1e6,
           "bandwidth_bound_ms": bytes_step / 6454.6e6, "note": "working set per layer < L2: latency / launch bound"}
    if with_cpu:
        from oracle import port
        torch.set_num_threads(os.cpu_count() or 1)
        torch.manual_seed(0)
        ref = port.OracleGCNModel(**BOTNET)
        xc, ei, yc = torch.from_numpy(g["x"]), torch.from_numpy(g["edge_index"]), torch.from_numpy(g["y"]).long()

        def cpu_step():
            ref.zero_grad()
            torch.nn.CrossEntropyLoss()(ref(xc[:, 0:1], ei, xc[:, 1]), yc).backward()

        cms = _cpu_time(cpu_step)
        out["cpu_baseline"] = {"ms": cms, "graphs_per_s": 1e3 / cms, "gedges_per_s": e * 24 / cms / 1e6,
                               "cores": torch.get_num_threads(), "kind": "port",
                               "sample": "the same graph, best of 2 steps after 1 warm-up"}
    return [out]


def config2(with_cpu=True):
    """configs[2]: kernel/gcn.py, gin.py (and graph_sage.py) 3 layers hidden 64 on TU-shaped batches of 128 graphs"""
    import torch
    from meta_gcn_b200 import data as D, kernel as K
    from meta_gcn_b200.graph import clear_structure_cache
    dev = torch.device("cuda")
    host = D.synth_tu_batch(seed=0, num_graphs=128)
    tb = host.to(dev)
    meta = D.dataset_meta(3, 2)
    res = []
    for name, cls, kind in (("GCN", K.GCN, "gcn"), ("GIN0", K.GIN0, "gin0"), ("GIN", K.GIN, "gin"),
                            ("GraphSAGE", K.GraphSAGE, "sage")):
        torch.manual_seed(0)
        net = cls(meta, 3, 64).to(dev).train()
        yb = tb.y.view(-1).long()

        def step():
            clear_structure_cache()          # a new mini-batch every step: structure build included
            net.zero_grad(set_to_none=True)
            torch.nn.functional.nll_loss(net(tb), yb).backward()

        ms = _timeit(step, reps=20, warm=5)
        r = {"config": 2, "what": f"configs[2]: kernel/{name} 3 layers hidden 64, batch of 128 TU-shaped graphs, fwd+bwd "
                                  f"incl. structure build", "net": name, "n": tb.num_nodes, "e": tb.num_edges, "ms": ms,
             "graphs_per_s": 128e3 / ms}
        if with_cpu:
            from oracle.kernel_nets import OracleGraphNet
            torch.set_num_threads(os.cpu_count() or 1)
            torch.manual_seed(0)
            ref = OracleGraphNet(kind, 3, 2, 3, 64).train()
            d = _Data()
            d.x, d.edge_index, d.batch = host.x, host.edge_index, host.batch
            yc = host.y.view(-1).long()

            def cpu_step():
                ref.zero_grad()
                torch.nn.functional.nll_loss(ref(d), yc).backward()

            cms = _cpu_time(cpu_step, reps=5, warm=2)
            r["cpu_baseline"] = {"ms": cms, "graphs_per_s": 128e3 / cms, "cores": torch.get_num_threads(),
                                 "kind": "port", "sample": "the same batch, best of 5 steps after 2 warm-ups"}
        res.append(r)
    return res


def config3(with_cpu=True):
    """configs[3]: kernel/graph_sage.py hidden 256 on an ogbn-products-shaped graph (tcgen05 wide transform)"""
    import torch
    from meta_gcn_b200 import data as D, kernel as K
    dev = torch.device("cuda")
    N4, E4 = 2_449_029, 61_859_140
    ei = torch.from_numpy(D.synth_powerlaw_graph(4, N4, E4, alpha=0.6)).to(dev)
    x = torch.randn(N4, 100, device=dev)
    yb = torch.randint(0, 47, (1,), device=dev)
    batch = D.GraphBatch(x, ei, yb, torch.zeros(N4, dtype=torch.long, device=dev), [0, N4])
    torch.manual_seed(0)
    net = K.GraphSAGE(D.dataset_meta(100, 47), 3, 256).to(dev).train()

    def fwd():
        with torch.no_grad():
            return net(batch)

    def step():
        net.zero_grad(set_to_none=True)
        torch.nn.functional.nll_loss(net(batch), yb).backward()

    ms_f = _timeit(fwd, reps=3, warm=1)
    ms_s = _timeit(step, reps=3, warm=1)
    r = {"config": 3, "what": "configs[3]: kernel/GraphSAGE 3 layers hidden 256 on a products-shaped graph (structures cached)",
         "n": N4, "e": E4, "fwd_ms": ms_f, "fwd_bwd_ms": ms_s, "gedges_per_s_fwd_bwd": E4 * 6 / ms_s / 1e6,
         "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
    if with_cpu:
        # the reference materialises [E,256] messages twice per layer (63 GB each at this shape): timed on a 1/16
        # sample of the shape and normalised to edges/s
        from oracle.kernel_nets import OracleGraphNet
        torch.set_num_threads(os.cpu_count() or 1)
        n_s, e_s = N4 // 16, E4 // 16
        eis = torch.from_numpy(D.synth_powerlaw_graph(4, n_s, e_s, alpha=0.6))
        d = _Data()
        d.x, d.edge_index, d.batch = torch.randn(n_s, 100), eis, torch.zeros(n_s, dtype=torch.long)
        torch.manual_seed(0)
        ref = OracleGraphNet("sage", 100, 47, 3, 256).train()
        yc = torch.randint(0, 47, (1,))

        def cpu_step():
            ref.zero_grad()
            torch.nn.functional.nll_loss(ref(d), yc).backward()

        cms = _cpu_time(cpu_step, reps=1, warm=1)
        r["cpu_baseline"] = {"ms": cms, "gedges_per_s_fwd_bwd": e_s * 6 / cms / 1e6, "cores": torch.get_num_threads(),
                             "kind": "port", "sample": f"1/16 of the shape (N={n_s}, E={e_s}), one fwd+bwd step after 1 warm-up"}
    return [r]


def config4(with_cpu=True):
    """configs[4]: aggregation sweep over average degree and width on power-law graphs vs the reference's
    index_select -> mul -> scatter_add path (CPU: oracle port; GPU: the same op sequence with stock torch CUDA ops)"""
    import torch
    from meta_gcn_b200 import data as D, ops
    from meta_gcn_b200.graph import GraphStructure, clear_structure_cache
    dev = torch.device("cuda")
    N5 = 1_000_000
    res = []
    for avg_deg in (2, 8, 32, 128):
        ei_h = torch.from_numpy(D.synth_powerlaw_graph(5, N5, N5 * avg_deg))
        ei = ei_h.to(dev)
        gs = GraphStructure(ei, N5)
        gs.fwd
        norm = torch.rand(ei.size(1), device=dev)
        for H in (16, 32, 64, 128, 256, 512):
            x = torch.randn(N5, H, device=dev)
            ours = _timeit(lambda: ops.spmm_impl(gs.fwd, x), reps=3, warm=1)
            plain = _timeit(lambda: ops.aggregate_prescaled_impl(gs.fwd, x), reps=3, warm=1) if H in (16, 32, 64, 128) else None

            def torch_path():
                xj = x.index_select(0, ei[0]) * norm.view(-1, 1)
                return torch.zeros(N5, H, device=dev).index_add_(0, ei[1], xj)

            ref = _timeit(torch_path, reps=2, warm=1) if ei.size(1) * H * 4 < 40e9 else None
            best = min(ours, plain or ours)
            r = {"config": 4, "avg_degree": avg_deg, "H": H, "n": N5, "e": int(ei.size(1)), "spmm_ms": ours,
                 "plain_ms": plain, "torch_index_add_ms": ref, "gather_gbs": ei.size(1) * H * 4 / best / 1e6,
                 "gedges_per_s": ei.size(1) / best / 1e6}
            if with_cpu and avg_deg in (8, 32) and H in (32, 256):
                from oracle import port
                torch.set_num_threads(os.cpu_count() or 1)
                xc, nc = x.cpu(), norm.cpu()
                cms = _cpu_time(lambda: port.scatter_rows("add", xc[ei_h[0]] * nc.view(-1, 1), ei_h[1], N5), reps=1, warm=1)
                r["cpu_baseline"] = {"ms": cms, "gedges_per_s": ei.size(1) / cms / 1e6, "cores": torch.get_num_threads(),
                                     "kind": "port", "sample": "the same aggregation (forward), one run after 1 warm-up"}
            res.append(r)
            del x
        del gs, ei, norm
        clear_structure_cache()
        torch.cuda.empty_cache()
    return res


CONFIGS = {0: config0, 2: config2, 3: config3, 4: config4}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="0,2,4,3")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    for k in [int(v) for v in a.only.split(",") if v.strip()]:
        for row in CONFIGS[k](not a.no_cpu):
            print(json.dumps(row), flush=True)
