"""ORACLE / TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the UNMODIFIED
reference (/root/reference, present only in the build container) over oracle/shim on CPU.

    python oracle/make_golden.py            # rewrites tests/golden/

The reference's own files that are executed:
  src/gcn_meta/models/{common,gcn_base_models,gcn_multi_kernel,gcn_model}.py   (botnet GCN path)
  src/gcn_meta/models/gcn.py                                                   (legacy GCN.norm)
  kernel/{gcn,gin,graph_sage}.py                                               (over the PyG restatement)
  data_procs/{undirected,loop}.py                                              (edge ordering)
Inputs are seeded; every case stores inputs, parameters, outputs, loss and parameter gradients.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

sys.path.insert(0, os.path.join(HERE, "shim"))
sys.path.insert(1, os.path.join(REF, "src"))
sys.path.insert(2, ROOT)


def load_file(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def small_graph(seed, n, m, loops=True):
    """undirected, sort-unique, self-loops appended — through the reference's own data_procs code"""
    und = load_file("ref_undirected", os.path.join(REF, "data_procs", "undirected.py"))
    loop = load_file("ref_loop", os.path.join(REF, "data_procs", "loop.py"))
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n, (m,), generator=g)
    dst = torch.randint(0, n, (m,), generator=g)
    keep = src != dst
    raw = torch.stack([src[keep], dst[keep]])
    ei, _ = und.to_undirected_ey(raw, None, n)
    if loops:
        ei, _ = loop.add_self_loops_ey(ei, None, None, n)
    return raw, ei


def np_state(model):
    return {"param." + k: v.detach().numpy().copy() for k, v in model.state_dict().items()}


def np_grads(model):
    return {"grad." + k: p.grad.detach().numpy().copy() for k, p in model.named_parameters()
            if p.grad is not None}


def case_gcn_meta(name, seed, n, m, model_kwargs, use_deg=True, edge_weight=False, graph_slices=None):
    from gcn_meta.models.gcn_model import GCNModel
    raw, ei = small_graph(seed, n, m)
    torch.manual_seed(seed)
    model = GCNModel(**model_kwargs)
    model.train()
    in_c = model_kwargs["in_channels"]
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.ones(n, 1) if in_c == 1 else torch.randn(n, in_c, generator=g)
    deg = torch.bincount(ei[0], minlength=n).float()
    ew = torch.rand(ei.size(1), generator=g) + 0.5 if edge_weight else None
    kw = {}
    if graph_slices is not None:
        kw["batch_slices_x"] = graph_slices
    out = model(x, ei, deg_K=deg if use_deg else None, edge_weight_K=ew, **kw)
    ncls = out.size(1)
    y = torch.randint(0, ncls, (out.size(0),), generator=g)
    loss = torch.nn.CrossEntropyLoss()(out, y)
    loss.backward()
    d = {"x": x.numpy(), "edge_index": ei.numpy(), "deg": deg.numpy(), "y": y.numpy(),
         "out": out.detach().numpy(), "loss": np.array(loss.item(), dtype=np.float64)}
    if ew is not None:
        d["edge_weight"] = ew.numpy()
    if graph_slices is not None:
        d["batch_slices_x"] = np.array(graph_slices)
    d.update(np_state(model))
    d.update(np_grads(model))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    return d


def case_gcn_meta_multi(name, seed, n, m, model_kwargs, num_sets=1, edge_attr_dim=None):
    """GCNModel over K edge sets (GCNMultiKernel, gcn_multi_kernel.py:76-114) and / or with per-edge attributes
    (gcn_base_models.py:204-206,227): inputs are lists, one entry per edge set"""
    from gcn_meta.models.gcn_model import GCNModel
    eis = [small_graph(seed + 17 * k, n, m)[1] for k in range(num_sets)]
    torch.manual_seed(seed)
    model = GCNModel(**model_kwargs)
    model.train()
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(n, model_kwargs["in_channels"], generator=g)
    degs = [torch.bincount(ei[0], minlength=n).float() for ei in eis]
    eas = [torch.randn(ei.size(1), edge_attr_dim, generator=g) for ei in eis] if edge_attr_dim else None
    out = model(x, eis if num_sets > 1 else eis[0], edge_attr_K=(eas if num_sets > 1 else eas[0]) if eas else None,
                deg_K=degs if num_sets > 1 else degs[0])
    y = torch.randint(0, out.size(1), (out.size(0),), generator=g)
    loss = torch.nn.CrossEntropyLoss()(out, y)
    loss.backward()
    d = {"x": x.numpy(), "y": y.numpy(), "out": out.detach().numpy(), "loss": np.array(loss.item(), dtype=np.float64),
         "num_sets": np.array(num_sets)}
    for k in range(num_sets):
        d[f"edge_index{k}"] = eis[k].numpy()
        d[f"deg{k}"] = degs[k].numpy()
        if eas:
            d[f"edge_attr{k}"] = eas[k].numpy()
    d.update(np_state(model))
    d.update(np_grads(model))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    return d


def case_primitives():
    """degnorm_const, NodeModelAdditive.forward, scatter_ and the legacy GCN.norm on one graph"""
    from gcn_meta.models.common import scatter_
    from gcn_meta.models.gcn_base_models import NodeModelAdditive, NodeModelBase
    legacy = load_file("ref_legacy_gcn", os.path.join(REF, "src", "gcn_meta", "models", "gcn.py"))
    n, m = 200, 900
    raw, ei = small_graph(7, n, m)
    g = torch.Generator().manual_seed(8)
    deg = torch.bincount(ei[0], minlength=n).float()
    ew = torch.rand(ei.size(1), generator=g) + 0.5
    d = {"edge_index": ei.numpy(), "raw_edge_index": raw.numpy(), "deg": deg.numpy(), "edge_weight": ew.numpy()}
    d["norm_sm"] = NodeModelBase.degnorm_const(ei, n, deg=deg, method="sm").numpy()
    d["norm_sm_nodeg"] = NodeModelBase.degnorm_const(ei, n, method="sm").numpy()
    d["norm_rw"] = NodeModelBase.degnorm_const(ei, n, deg=deg, method="rw").numpy()
    d["norm_sm_w"] = NodeModelBase.degnorm_const(ei, n, edge_weight=ew, method="sm").numpy()
    d["norm_rw_w"] = NodeModelBase.degnorm_const(ei, n, edge_weight=ew, method="rw").numpy()
    # zero-degree handling: a graph whose last 5 nodes have no out-edges at all
    ei_iso = ei[:, (ei[0] < n - 5)]
    d["edge_index_iso"] = ei_iso.numpy()
    d["norm_sm_iso"] = NodeModelBase.degnorm_const(ei_iso, n, method="sm").numpy()
    x = torch.randn(n, 16, generator=g)
    d["x"] = x.numpy()
    for aggr in ("add", "mean"):
        for dn in ("sm", "rw", None):
            torch.manual_seed(3)
            nm = NodeModelAdditive(16, 32, deg_norm=dn, aggr=aggr, bias=True)
            with torch.no_grad():
                nm.bias.uniform_(-0.1, 0.1)
            tag = f"additive_{aggr}_{dn}"
            d[tag + ".weight_node"] = nm.weight_node.detach().numpy().copy()
            d[tag + ".bias"] = nm.bias.detach().numpy().copy()
            d[tag + ".out"] = nm(x, ei, deg=deg).detach().numpy()
    src = torch.randn(ei.size(1), 8, generator=g)
    d["scatter_src"] = src.numpy()
    d["scatter_add"] = scatter_("add", src, ei[1], dim_size=n).numpy()
    d["scatter_mean"] = scatter_("mean", src, ei[1], dim_size=n).numpy()
    # legacy operator's norm(): appends loops itself -> run on the loop-free undirected graph
    _, ei_noloop = small_graph(7, n, m, loops=False)
    li, ln = legacy.GCN.norm(ei_noloop, n, None, dtype=torch.float32)
    d["legacy_in_edge_index"] = ei_noloop.numpy()
    d["legacy_edge_index"] = li.numpy()
    d["legacy_norm"] = ln.numpy()
    np.savez_compressed(os.path.join(OUT, "primitives.npz"), **d)


def tu_batch(seed, num_graphs, f_in):
    from meta_gcn_b200.data import synth_tu_batch
    return synth_tu_batch(seed, num_graphs, f_in, 2)


class _DS:
    def __init__(self, f, c):
        self.num_features, self.num_classes = f, c


def case_kernel_net(name, modfile, cls, seed, num_layers=3, hidden=64, **extra):
    mod = load_file("ref_kernel_" + modfile, os.path.join(REF, "kernel", modfile + ".py"))
    b = tu_batch(seed, 24, 3)
    torch.manual_seed(seed)
    model = getattr(mod, cls)(_DS(3, 2), num_layers, hidden, **extra)
    state0 = np_state(model)  # before any forward: BatchNorm running stats still (0, 1)
    model.eval()  # dropout off; BatchNorm uses running stats -> a train-mode pass is stored as well
    out_eval = model(b).detach().numpy()
    model.train()
    torch.manual_seed(seed + 100)  # F.dropout stream
    # dropout(p=0.5) draws from the global RNG; gradients are taken with dropout disabled instead
    import torch.nn.functional as F
    real_dropout = F.dropout
    F.dropout = lambda x, p=0.5, training=True, inplace=False: x
    try:
        out = model(b)
        loss = F.nll_loss(out, b.y.view(-1))
        loss.backward()
    finally:
        F.dropout = real_dropout
    d = {"x": b.x.numpy(), "edge_index": b.edge_index.numpy(), "batch": b.batch.numpy(), "y": b.y.numpy(),
         "out_eval": out_eval, "out_train_nodrop": out.detach().numpy(),
         "loss": np.array(loss.item(), dtype=np.float64)}
    d.update(state0)
    d.update(np_grads(model))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)


def case_preprocess():
    n, m = 500, 3000
    raw, ei = small_graph(21, n, m)
    deg = torch.bincount(ei[0], minlength=n).float()
    np.savez_compressed(os.path.join(OUT, "preprocess.npz"), raw=raw.numpy(), edge_index=ei.numpy(),
                        deg=deg.numpy(), num_nodes=np.array(n))


def case_metrics():
    """binary metrics of src/gcn_meta/optim/metrics.py on seeded logits / labels (train_botnet.py:296-305)"""
    met = load_file("ref_metrics", os.path.join(REF, "src", "gcn_meta", "optim", "metrics.py"))
    out = {}
    for k, (n, p1, seed) in enumerate([(5000, 0.07, 0), (300, 0.5, 1), (64, 0.0, 2)]):
        g = torch.Generator().manual_seed(seed)
        logits = torch.randn(n, 2, generator=g)
        y = (torch.rand(n, generator=g) < p1).long()
        logits[:, 1] += 2.0 * y.float() - 1.0
        pred = logits.max(1)[1]
        vals = [met.accuracy(pred, y), met.true_positive(pred, y), met.false_positive(pred, y),
                met.true_negative(pred, y), met.false_negative(pred, y)]
        for fn in (met.recall, met.precision, met.f1_score, met.false_positive_rate, met.false_negative_rate):
            try:
                vals.append(float(fn(pred, y)))
            except ZeroDivisionError:
                vals.append(float("nan"))
        out[f"logits{k}"] = logits.numpy()
        out[f"y{k}"] = y.numpy()
        out[f"vals{k}"] = np.array(vals, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), **out)


def case_primitive_max():
    """scatter_('max', src, index, dim_size) of common.py:37-66 (fill -1e38, untouched rows -> 0) and its autograd
    (torch_scatter scatter_max: gradient to the first maximal entry), incl. rows without entries and ties"""
    from gcn_meta.models.common import scatter_
    n, m = 120, 700
    _, ei = small_graph(31, n, m, loops=False)
    ei = ei[:, ei[1] < n - 7]                      # the last 7 rows receive nothing
    g = torch.Generator().manual_seed(31)
    src = torch.randn(ei.size(1), 8, generator=g)
    src[5] = src[9] = src[2]                       # exact ties
    src.requires_grad_(True)
    wout = torch.randn(n, 8, generator=g)
    out = scatter_("max", src, ei[1], dim_size=n)
    (out * wout).sum().backward()
    np.savez_compressed(os.path.join(OUT, "primitive_max.npz"), index=ei[1].numpy(), src=src.detach().numpy(),
                        wout=wout.numpy(), out=out.detach().numpy(), grad_src=src.grad.numpy(), num_nodes=np.array(n))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    if len(sys.argv) > 1 and sys.argv[1] == "max":   # only the 'max' aggregation cases (added later)
        v_ = dict(in_channels=1, enc_sizes=[16, 16, 16], num_classes=2, non_linear="relu",
                  non_linear_layer_wise="relu", residual_hop=1, dropout=0.0, final_type="proj", pred_on="node",
                  nodemodel="additive", deg_norm="sm", edge_gate=None, aggr="max", bias=True)
        case_gcn_meta("gcn_meta_max", 7, 150, 500, v_)
        case_gcn_meta("gcn_meta_max_ew_rw", 8, 150, 500, dict(v_, in_channels=5, deg_norm="rw"), edge_weight=True)
        case_primitive_max()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "attention":   # soft-attention node model (added later)
        v_ = dict(in_channels=5, enc_sizes=[16, 16, 16], num_classes=2, non_linear="relu",
                  non_linear_layer_wise="relu", residual_hop=1, dropout=0.0, final_type="proj", pred_on="node",
                  nodemodel="attention", nheads=2, att_act="lrelu", att_dropout=0, att_combine="cat", att_dir="in",
                  bias=True)
        case_gcn_meta("gcn_meta_attention", 11, 150, 500, v_, use_deg=False)
        case_gcn_meta("gcn_meta_attention_out_mean", 12, 150, 500,
                      dict(v_, nheads=[2, 4, 1], att_combine="mean", att_dir="out", att_act="relu"), use_deg=False)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "r2":    # round 2: per-layer deg_norm override, edge attributes, two edge sets
        v_ = dict(in_channels=5, enc_sizes=[16, 16, 16], num_classes=2, non_linear="relu",
                  non_linear_layer_wise="relu", residual_hop=1, dropout=0.0, final_type="proj", pred_on="node",
                  nodemodel="additive", deg_norm="sm", edge_gate=None, aggr="add", bias=True)
        case_gcn_meta("gcn_meta_final_rw", 21, 150, 500, dict(v_, final_layer_config={"deg_norm": "rw"}))
        case_gcn_meta("gcn_meta_final_rw_nobias32", 22, 300, 1200,
                      dict(v_, in_channels=1, enc_sizes=[32] * 4, bias=False, final_layer_config={"deg_norm": "rw"}))
        case_gcn_meta_multi("gcn_meta_edgeattr", 23, 150, 500, dict(v_, in_edgedim=3), edge_attr_dim=3)
        case_gcn_meta_multi("gcn_meta_two_kernels_add", 24, 150, 400, dict(v_, num_kernel=2, kernel_combine="add"),
                            num_sets=2)
        case_gcn_meta_multi("gcn_meta_max_edgeattr", 25, 150, 500, dict(v_, aggr="max", in_edgedim=3), edge_attr_dim=3)
        case_gcn_meta("gcn_meta_max_gate_proj", 26, 150, 500, dict(v_, aggr="max", edge_gate="proj"))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "gate":  # edge-gate cases (added later)
        v_ = dict(in_channels=1, enc_sizes=[16, 16, 16], num_classes=2, non_linear="relu",
                  non_linear_layer_wise="relu", residual_hop=1, dropout=0.0, final_type="proj", pred_on="node",
                  nodemodel="additive", deg_norm="sm", edge_gate="proj", aggr="add", bias=True)
        case_gcn_meta("gcn_meta_gate_proj", 9, 150, 500, v_)
        case_gcn_meta("gcn_meta_gate_proj_mean_ew", 10, 150, 500, dict(v_, in_channels=5, aggr="mean", deg_norm="rw"),
                      edge_weight=True)
        return
    botnet = dict(in_channels=1, enc_sizes=[32] * 12, num_classes=2, non_linear="relu",
                  non_linear_layer_wise="relu", residual_hop=1, dropout=0.0, final_type="proj",
                  pred_on="node", nodemodel="additive", deg_norm="sm", edge_gate=None, aggr="add",
                  bias=False)
    case_gcn_meta("gcn_meta_botnet12", 0, 400, 1600, botnet)
    v = dict(botnet, enc_sizes=[16, 16, 16], bias=True)
    case_gcn_meta("gcn_meta_rw_bias", 1, 150, 500, dict(v, deg_norm="rw"))
    case_gcn_meta("gcn_meta_nonorm_mean", 2, 150, 500, dict(v, deg_norm=None, aggr="mean"))
    case_gcn_meta("gcn_meta_hop2_none", 3, 150, 500,
                  dict(v, enc_sizes=[16, 16, 16, 16], residual_hop=2, final_type="none", num_classes=16))
    case_gcn_meta("gcn_meta_edgeweight", 4, 150, 500, dict(v, in_channels=5), edge_weight=True)
    case_gcn_meta("gcn_meta_nodeg", 5, 150, 500, dict(v, in_channels=5), use_deg=False)
    case_gcn_meta("gcn_meta_graphpred", 6, 160, 500, dict(v, in_channels=5, pred_on="graph"),
                  graph_slices=[0, 160])  # multi-graph batches hit a shape bug in gcn_model.py:123 unless B == C
    case_primitives()
    case_preprocess()
    case_metrics()
    case_kernel_net("kernel_gcn", "gcn", "GCN", 10)
    case_kernel_net("kernel_gcn_jk", "gcn", "GCNWithJK", 11)
    case_kernel_net("kernel_gin0", "gin", "GIN0", 12)
    case_kernel_net("kernel_gin", "gin", "GIN", 13)
    case_kernel_net("kernel_sage", "graph_sage", "GraphSAGE", 14)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
