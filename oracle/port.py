"""ORACLE / TEST INFRASTRUCTURE ONLY — never imported by the product path (meta_gcn_b200/).

CPU restatement (plain torch CPU ops + numpy) of the reference's message-passing hot path, each
function citing the reference lines it follows.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module.

Parity status: PINNED for the src/gcn_meta path — oracle/make_golden.py runs the unmodified
reference modules (imported from /root/reference over oracle/shim) and stores their outputs in
tests/golden/; tests/test_oracle_golden.py checks this restatement against them.  The PyG-operator
path (kernel/gcn.py, gin.py, graph_sage.py) is anchored on a restatement of the un-vendored
torch_geometric 1.3 operators (oracle/shim/torch_geometric) — "parity unpinned" for those beyond the
in-repo legacy GCN operator (src/gcn_meta/models/gcn.py:57-107), which IS executed for the goldens.
"""
import math

import numpy as np
import torch
import torch.nn as nn


# ------------------------------------------------------------------------------------------------
# primitives
# ------------------------------------------------------------------------------------------------
def scatter_rows(name, src, index, dim_size):
    """common.py:37-66 -> torch_scatter 1.x: out.scatter_add_(0, index.expand_as(src), src);
    'mean' divides by the clamped count.  CPU scatter_add_ sums in edge order (deterministic)."""
    assert name in ("add", "mean", "max")
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    if name == "max":
        # common.py:56-64: scatter_max from fill -1e38, untouched rows -> 0.  amax's autograd splits the gradient
        # between tied maxima; torch_scatter gives it to the FIRST one — equal forward values, and the gradient
        # differs only on exact ties (the golden cases carry ties only in the primitive test, compared forward-only
        # through this function and with the explicit first-argmax rule in segment_max_first below)
        fill = -1e38
        out = src.new_full((dim_size,) + tuple(src.shape[1:]), fill).scatter_reduce(0, idx, src, "amax", include_self=True)
        return torch.where(out == fill, torch.zeros_like(out), out)
    out = src.new_zeros((dim_size,) + tuple(src.shape[1:])).scatter_add_(0, idx, src)
    if name == "mean":
        cnt = src.new_zeros((dim_size,) + tuple(src.shape[1:])).scatter_add_(0, idx, torch.ones_like(src))
        out = out / cnt.clamp(min=1)
    return out


def segment_max_first(src, index, dim_size):
    """torch_scatter 1.x scatter_max restated with numpy loops: (out, arg) with arg = the FIRST entry attaining the
    maximum, -1 / fill for untouched rows; its backward sends grad_out[i,c] to src[arg[i,c], c]."""
    s, ix = src.detach().numpy(), index.numpy()
    out = np.full((dim_size,) + s.shape[1:], -1e38, dtype=s.dtype)
    arg = np.full(out.shape, -1, dtype=np.int64)
    for e in range(s.shape[0]):
        better = s[e] > out[ix[e]]
        out[ix[e]][better] = s[e][better]
        arg[ix[e]][better] = e
    return out, arg


def degnorm_const(edge_index, num_nodes, deg=None, edge_weight=None, method="sm"):
    """gcn_base_models.py:65-146.  Returns norm[E] ('sm', or 'rw' with weights) or dis[N] ('rw')."""
    assert method in ("sm", "rw")
    row, col = edge_index
    equal = edge_weight is None
    if not equal:
        edge_weight = edge_weight.view(-1)
        deg = scatter_rows("add", edge_weight, row, num_nodes)                      # :126
    elif deg is None:
        deg = scatter_rows("add", torch.ones(edge_index.size(1), dtype=torch.float32), row, num_nodes)
    dis = deg.pow(-0.5) if method == "sm" else deg.pow(-1)                          # :128-131
    dis = dis.clone()
    dis[dis == float("inf")] = 0                                                    # :135
    if method == "sm":                                                              # :138-139
        return dis[row] * dis[col] if equal else dis[row] * edge_weight * dis[col]
    return dis if equal else dis[row] * edge_weight                                # :141-142


def additive_node_model(x, edge_index, weight_node, bias=None, deg=None, edge_weight=None,
                        deg_norm="sm", aggr="add"):
    """NodeModelAdditive.forward, gcn_base_models.py:199-243 (no edge features / gates)."""
    x = torch.matmul(x, weight_node)                                                # :201
    if deg_norm is None:
        x_j = torch.index_select(x, 0, edge_index[0])                               # :211
    else:
        norm = degnorm_const(edge_index, x.size(0), deg, edge_weight, deg_norm)     # :215
        if deg_norm == "rw" and edge_weight is None:
            x_j = torch.index_select(x * norm.view(-1, 1), 0, edge_index[0])        # :218-220
        else:
            x_j = torch.index_select(x, 0, edge_index[0]) * norm.view(-1, 1)        # :223-224
    out = scatter_rows(aggr, x_j, edge_index[1], x.size(0))                         # :237
    if bias is not None:
        out = out + bias                                                            # :240
    return out


# ------------------------------------------------------------------------------------------------
# the botnet model: GCNModel / GCNLayer / GCNMultiKernel (K=1) / NodeModelAdditive
# ------------------------------------------------------------------------------------------------
def _glorot_(t):
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-a, a)


class _NodeModel(nn.Module):
    def __init__(self, cin, cout, bias):
        super().__init__()
        self.weight_node = nn.Parameter(torch.empty(cin, cout))
        if bias:
            self.bias = nn.Parameter(torch.empty(cout))
        else:
            self.register_parameter("bias", None)
        _glorot_(self.weight_node)                                                  # :193
        if bias:
            nn.init.zeros_(self.bias)


class _MultiKernel(nn.Module):
    def __init__(self, cin, cout, bias):
        super().__init__()
        self.node_models = nn.ModuleList([_NodeModel(cin, cout, bias)])


class _Layer(nn.Module):
    def __init__(self, cin, cout, bias):
        super().__init__()
        self.gcn = _MultiKernel(cin, cout, bias)


class OracleGCNModel(nn.Module):
    """Restates GCNModel (gcn_model.py:8-125) for the additive node model, num_kernel=1, with the
    reference's parameter names (gcn_net.N.gcn.node_models.0.weight_node, residuals.N.{weight,bias},
    final.{weight,bias}) and creation order, so the same torch seed gives the same weights."""

    def __init__(self, in_channels, enc_sizes, num_classes, non_linear="relu",
                 non_linear_layer_wise="relu", residual_hop=None, dropout=0.5, final_type="none",
                 pred_on="node", deg_norm="sm", aggr="add", bias=True, **unused):
        super().__init__()
        self.sizes = [in_channels, *enc_sizes]
        self.num_layers = len(enc_sizes)
        self.residual_hop = residual_hop
        self.deg_norm, self.aggr = deg_norm, aggr
        self.pred_on = pred_on
        self.act_layer = non_linear_layer_wise
        self.act_res = non_linear
        self.gcn_net = nn.ModuleList([_Layer(a, b, bias) for a, b in zip(self.sizes, self.sizes[1:])])
        self.dropout = nn.Dropout(dropout)
        if residual_hop is not None and residual_hop > 0:                           # :61-67
            self.residuals = nn.ModuleList([
                nn.Linear(self.sizes[i], self.sizes[j])
                for i, j in zip(range(0, len(self.sizes), residual_hop),
                                range(residual_hop, len(self.sizes), residual_hop))])
        self.final = nn.Linear(self.sizes[-1], num_classes) if final_type == "proj" else nn.Identity()

    @staticmethod
    def _act(name, t):
        if name == "relu":
            return torch.relu(t)
        if name == "none":
            return t
        if name == "lrelu":
            return nn.functional.leaky_relu(t, 0.2)
        if name == "elu":
            return nn.functional.elu(t)
        raise ValueError(name)

    def forward(self, x, edge_index, deg=None, edge_weight=None, batch_slices_x=None):
        xr, add_at = None, -1
        hop = self.residual_hop
        for n, layer in enumerate(self.gcn_net):                                    # :89
            nm = layer.gcn.node_models[0]
            xo = additive_node_model(x, edge_index, nm.weight_node, nm.bias, deg, edge_weight,
                                     self.deg_norm, self.aggr)
            xo = self._act(self.act_layer, xo)                                      # :196
            xo = self.dropout(xo)                                                   # :92
            if hop is not None and hop > 0:
                if n % hop == 0 and (n // hop) < len(self.residuals):               # :95-97
                    xr = self.residuals[n // hop](x)
                    add_at = n + hop - 1
                if n == add_at:
                    xo = self._act(self.act_res, xo + xr) if n < self.num_layers - 1 else xo + xr
            x = xo
        x = self.final(x)                                                           # :108
        if self.pred_on == "graph":                                                 # :112-123
            sl = batch_slices_x
            x = torch.stack([x[i:j].sum(0) / (j - i) for i, j in zip(sl, sl[1:])])
        return x


# ------------------------------------------------------------------------------------------------
# offline edge preprocessing that defines the edge ORDER (data_procs/)
# ------------------------------------------------------------------------------------------------
def sort_unique_edges(edge_index, num_nodes):
    """data_procs/undirected.py:6-16: unique(row*N+col) -> lexicographic (src,dst), duplicates
    dropped; perm = index of one occurrence of each kept key (torch_sparse picks the first)."""
    row, col = np.asarray(edge_index[0]), np.asarray(edge_index[1])
    key = row.astype(np.int64) * num_nodes + col
    _, perm = np.unique(key, return_index=True)
    return np.stack([row[perm], col[perm]]), perm


def to_undirected(edge_index, num_nodes):
    """data_procs/undirected.py:19-35."""
    row, col = np.asarray(edge_index[0]), np.asarray(edge_index[1])
    both = np.stack([np.concatenate([row, col]), np.concatenate([col, row])])
    return sort_unique_edges(both, num_nodes)[0]


def append_self_loops(edge_index, num_nodes):
    """data_procs/loop.py:13-17: (i,i) for every node appended at the END."""
    loop = np.arange(num_nodes, dtype=np.int64)
    return np.concatenate([np.asarray(edge_index), np.stack([loop, loop])], axis=1)


def out_degree(edge_index, num_nodes):
    """data_procs/data_add_degree.py:45-65: scatter_add(ones, row) as float32."""
    return np.bincount(np.asarray(edge_index[0]), minlength=num_nodes).astype(np.float32)


# ------------------------------------------------------------------------------------------------
# integer oracles for the structure build
# ------------------------------------------------------------------------------------------------
def csr_oracle(edge_index, num_nodes, by, loop_mode=0):
    """Stable grouping of edge_index by endpoint `by` (0 source / 1 target) with PyG loop handling
    (Appendix A: remove_self_loops / add_remaining_self_loops append loops at the END).
    Returns rowptr[N+1], nbr[nnz], perm[nnz] (perm >= E marks appended loop of node perm-E)."""
    ei = np.asarray(edge_index, dtype=np.int64)
    E = ei.shape[1]
    pos = np.arange(E, dtype=np.int64)
    src, dst = ei[0], ei[1]
    if loop_mode in (1, 2):
        keep = src != dst
        src, dst, pos = src[keep], dst[keep], pos[keep]
    if loop_mode == 2:
        loop = np.arange(num_nodes, dtype=np.int64)
        src, dst = np.concatenate([src, loop]), np.concatenate([dst, loop])
        pos = np.concatenate([pos, E + loop])
    key, other = (src, dst) if by == 0 else (dst, src)
    order = np.argsort(key, kind="stable")
    rowptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(np.bincount(key, minlength=num_nodes), out=rowptr[1:])
    return rowptr.astype(np.int32), other[order].astype(np.int32), pos[order].astype(np.int32)


def aggregate_dense_f64(edge_index, num_nodes, x, weights=None):
    """Independent fp64 arbiter: out = A_w^T-convention dense matmul, A[t,s] += w_e."""
    ei = np.asarray(edge_index)
    A = np.zeros((num_nodes, num_nodes), dtype=np.float64)
    w = np.ones(ei.shape[1]) if weights is None else np.asarray(weights, dtype=np.float64)
    np.add.at(A, (ei[1], ei[0]), w)
    return A @ np.asarray(x, dtype=np.float64)


def binary_metrics(pred, target):
    """src/gcn_meta/optim/metrics.py:8-60 restated on numpy integer arrays: returns
    [accuracy, TP, FP, TN, FN, recall, precision, f1, fpr, fnr] with NaN where the reference raises
    ZeroDivisionError (recall / fpr / fnr, and f1 through recall), -1 / 0 where it catches it
    (precision :38-39, f1 :47-48)."""
    pred = np.asarray(pred).astype(np.int64)
    target = np.asarray(target).astype(np.int64)
    tp = int(((pred == 1) & (target == 1)).sum())      # :13
    fp = int(((pred == 1) & (target == 0)).sum())      # :17
    tn = int(((pred == 0) & (target == 0)).sum())      # :21
    fn = int(((pred == 0) & (target == 1)).sum())      # :25
    acc = int((pred == target).sum()) / target.size    # :9
    pos, neg, ppos = int((target == 1).sum()), int((target == 0).sum()), int((pred == 1).sum())
    nan = float("nan")
    rec = tp / pos if pos else nan                     # :31
    prec = tp / ppos if ppos else -1                   # :35-39
    if pos == 0:
        f1 = nan                                       # recall raises before the try of :46
    elif prec + rec == 0:
        f1 = 0                                         # :47-48
    else:
        f1 = 2 * (prec * rec) / (prec + rec)
    fpr = fp / neg if neg else nan                     # :52
    fnr = fn / pos if pos else nan                     # :59
    return np.array([acc, tp, fp, tn, fn, rec, prec, f1, fpr, fnr], dtype=np.float64)
