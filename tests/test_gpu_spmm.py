"""Row-owned aggregation vs the oracle's index_select -> mul -> scatter_add (CPU, edge order)."""
import numpy as np
import pytest
import torch

from meta_gcn_b200 import functional as F_mgcn
from meta_gcn_b200 import ops
from meta_gcn_b200.graph import GraphStructure
from oracle import port
from util import assert_bitexact, assert_parity

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rand_graph(seed, n, e, hub=None):
    rng = np.random.default_rng(seed)
    ei = rng.integers(0, n, size=(2, e))
    if hub:
        ei[1, : hub] = 0          # node 0 receives `hub` messages
        ei[0, hub: 2 * hub] = 1   # node 1 sends `hub` messages
    return ei


def oracle_aggregate(ei, n, x, w=None, reduce="add"):
    ei_t = torch.from_numpy(ei)
    msg = x[ei_t[0]]
    if w is not None:
        msg = msg * w.view(-1, 1)
    return port.scatter_rows(reduce, msg, ei_t[1], n)


@pytest.mark.parametrize("H", [1, 2, 3, 4, 8, 12, 16, 32, 33, 64, 100, 128, 256, 512])
def test_plain_sum_is_bitexact_below_hub_threshold(H):
    n, e = 700, 9000
    ei = rand_graph(H, n, e)
    x = torch.randn(n, H)
    g = GraphStructure(torch.from_numpy(ei).to(DEV), n, hub_threshold=1 << 30)
    out = ops.spmm_impl(g.fwd, x.to(DEV))
    assert_bitexact(out, oracle_aggregate(ei, n, x), f"sum H={H}")
    outm = ops.spmm_impl(g.fwd, x.to(DEV), reduce=1)
    assert_bitexact(outm, oracle_aggregate(ei, n, x, reduce="mean"), f"mean H={H}")


@pytest.mark.parametrize("H", [1, 32, 64, 100, 256])
def test_sym_norm_weights_follow_reference_rounding(H):
    """w_e = dis[row]*dis[col] (gcn_base_models.py:138-139), msg = x_j * w_e, edge-order sum"""
    n, e = 900, 12000
    ei = rand_graph(100 + H, n, e)
    x = torch.randn(n, H)
    deg = torch.bincount(torch.from_numpy(ei[0]), minlength=n).float()
    norm = port.degnorm_const(torch.from_numpy(ei), n, deg=deg, method="sm")
    g = GraphStructure(torch.from_numpy(ei).to(DEV), n, hub_threshold=1 << 30)
    dis = ops.gcn_norm_impl(deg.to(DEV), 0)
    out = ops.spmm_impl(g.fwd, x.to(DEV), nbr_scale=dis, row_scale=dis)
    assert_bitexact(out, oracle_aggregate(ei, n, x, norm), f"sm H={H}")
    # with edge weights: dis[row] * w * dis[col]
    ew = torch.rand(e) + 0.5
    normw = port.degnorm_const(torch.from_numpy(ei), n, edge_weight=ew, method="sm")
    degw = g.weighted_out_degree(ew.to(DEV))
    disw = ops.gcn_norm_impl(degw, 0)
    ev_f, _ = g.edge_values(ew.to(DEV))
    outw = ops.spmm_impl(g.fwd, x.to(DEV), edge_val=ev_f, nbr_scale=disw, row_scale=disw)
    assert_bitexact(outw, oracle_aggregate(ei, n, x, normw), f"sm+w H={H}")
    # rw: x * deg^-1 gathered (gcn_base_models.py:217-220)
    dis_rw = ops.gcn_norm_impl(deg.to(DEV), 1)
    out_rw = ops.spmm_impl(g.fwd, x.to(DEV), nbr_scale=dis_rw)
    ref_rw = port.scatter_rows("add", (x * deg.pow(-1).nan_to_num(posinf=0).view(-1, 1))[torch.from_numpy(ei[0])],
                               torch.from_numpy(ei[1]), n)
    assert_bitexact(out_rw, ref_rw, f"rw H={H}")


@pytest.mark.parametrize("H", [3, 32, 64, 256])
@pytest.mark.parametrize("hub_t", [16, 256])
def test_hub_rows_tree_sum_within_tolerance_and_deterministic(H, hub_t):
    n, e = 500, 20000
    ei = rand_graph(7 + H, n, e, hub=6000)
    x = torch.randn(n, H)
    g = GraphStructure(torch.from_numpy(ei).to(DEV), n, hub_threshold=hub_t)
    assert int(g.fwd.hub_count.item()) >= 1
    out = ops.spmm_impl(g.fwd, x.to(DEV))
    ref = oracle_aggregate(ei, n, x)
    assert_parity(out, ref, f"hub sum H={H}")
    # non-hub rows stay bit-exact
    deg_in = np.bincount(ei[1], minlength=n)
    small = torch.from_numpy(deg_in <= hub_t)
    assert_bitexact(out.cpu()[small], ref[small], "non-hub rows")
    again = ops.spmm_impl(g.fwd, x.to(DEV))
    assert_bitexact(again, out, "run-to-run determinism")
    outm = ops.spmm_impl(g.fwd, x.to(DEV), reduce=1)
    assert_parity(outm, oracle_aggregate(ei, n, x, reduce="mean"), f"hub mean H={H}")


@pytest.mark.parametrize("H", [5, 32])
def test_fused_epilogue(H):
    n, e = 300, 3000
    ei = rand_graph(11, n, e)
    x, res, bias = torch.randn(n, H), torch.randn(n, H), torch.randn(H)
    g = GraphStructure(torch.from_numpy(ei).to(DEV), n, hub_threshold=1 << 30)
    out = ops.spmm_impl(g.fwd, x.to(DEV), bias=bias.to(DEV), residual=res.to(DEV), act=1)
    ref = torch.relu(oracle_aggregate(ei, n, x) + bias + res)
    assert_bitexact(out, ref, "relu(sum + bias + residual)")


def test_rows_without_edges_and_empty_inputs():
    n = 50
    ei = np.array([[1, 2, 3], [4, 4, 9]])
    x = torch.randn(n, 32)
    g = GraphStructure(torch.from_numpy(ei).to(DEV), n)
    out = ops.spmm_impl(g.fwd, x.to(DEV), reduce=1, bias=torch.ones(32, device=DEV))
    ref = oracle_aggregate(ei, n, x, reduce="mean") + 1
    assert_bitexact(out, ref, "isolated rows")
    g0 = GraphStructure(torch.zeros(2, 0, dtype=torch.long, device=DEV), n)
    out0 = ops.spmm_impl(g0.fwd, x.to(DEV))
    assert (out0 == 0).all()


def test_scatter_rows_primitive_seam():
    """scatter_('add'|'mean', src[E,H], index) — common.py:37-66 — incl. 1-D src and autograd"""
    n, e = 400, 6000
    rng = np.random.default_rng(3)
    idx = torch.from_numpy(rng.integers(0, n, size=e))
    for shape in ((e, 16), (e,), (e, 3)):
        src = torch.randn(*shape)
        for name in ("add", "mean"):
            out = F_mgcn.scatter_rows(src.to(DEV), idx.to(DEV), n, name)
            assert_bitexact(out, port.scatter_rows(name, src, idx, n), f"scatter_{name}{shape}")
    src = torch.randn(e, 8, requires_grad=True)
    src_d = src.detach().to(DEV).requires_grad_(True)
    wgt = torch.randn(n, 8)
    (F_mgcn.scatter_rows(src_d, idx.to(DEV), n, "mean") * wgt.to(DEV)).sum().backward()
    (port.scatter_rows("mean", src, idx, n) * wgt).sum().backward()
    assert_parity(src_d.grad, src.grad, "scatter_mean backward")


@pytest.mark.parametrize("reduce", ["add", "mean"])
@pytest.mark.parametrize("H", [32, 48])
def test_aggregate_backward_matches_autograd_of_reference_path(reduce, H):
    n, e = 600, 8000
    ei = rand_graph(21, n, e, hub=700)
    deg = torch.bincount(torch.from_numpy(ei[0]), minlength=n).float().clamp(min=1)
    ew = torch.rand(e) + 0.5
    x = torch.randn(n, H, requires_grad=True)
    res = torch.randn(n, H, requires_grad=True)
    bias = torch.randn(H, requires_grad=True)
    wgt = torch.randn(n, H)
    # reference path on CPU
    ei_t = torch.from_numpy(ei)
    dis = deg.pow(-0.5)
    norm = dis[ei_t[0]] * ew * dis[ei_t[1]]
    ref = torch.relu(port.scatter_rows(reduce, x[ei_t[0]] * norm.view(-1, 1), ei_t[1], n) + bias + res)
    (ref * wgt).sum().backward()
    # CUDA path
    xd, rd, bd = (t.detach().to(DEV).requires_grad_(True) for t in (x, res, bias))
    g = GraphStructure(ei_t.to(DEV), n, hub_threshold=64)
    disd = ops.gcn_norm_impl(deg.to(DEV), 0)
    out = F_mgcn.aggregate(xd, g, disd, disd, ew.to(DEV), reduce, bd, rd, "relu")
    (out * wgt.to(DEV)).sum().backward()
    assert_parity(out, ref, "forward")
    assert_parity(xd.grad, x.grad, "dx")
    assert_parity(rd.grad, res.grad, "dresidual")
    assert_parity(bd.grad, bias.grad, "dbias")


def test_backward_is_deterministic():
    n, e, H = 2000, 60000, 32
    ei = rand_graph(5, n, e, hub=3000)
    g = GraphStructure(torch.from_numpy(ei).to(DEV), n)
    x = torch.randn(n, H, device=DEV, requires_grad=True)
    wgt = torch.randn(n, H, device=DEV)
    grads = []
    for _ in range(3):
        x.grad = None
        (F_mgcn.aggregate(x, g) * wgt).sum().backward()
        grads.append(x.grad.clone())
    assert_bitexact(grads[1], grads[0], "dx run 2")
    assert_bitexact(grads[2], grads[0], "dx run 3")


@pytest.mark.parametrize("H", [16, 32, 64, 128])
@pytest.mark.parametrize("reduce", [0, 1])
def test_prescaled_aggregation(H, reduce):
    """inputs carrying the per-source factor: out = post * sum x~[nbr]  — same summation order as
    the exact kernel, different placement of the two roundings of the degree factors"""
    n, e = 3000, 40000
    ei = rand_graph(50 + H, n, e, hub=2500)
    deg = torch.bincount(torch.from_numpy(ei[0]), minlength=n).float().clamp(min=1)
    dis = deg.pow(-0.5)
    x = torch.randn(n, H)
    bias, res = torch.randn(H), torch.randn(n, H)
    g = GraphStructure(torch.from_numpy(ei).to(DEV), n, hub_threshold=64)
    xs = (x * dis.view(-1, 1)).to(DEV)
    out = ops.aggregate_prescaled_impl(g.fwd, xs, dis.to(DEV), reduce, bias.to(DEV), res.to(DEV), 1)
    norm = dis[torch.from_numpy(ei[0])] * dis[torch.from_numpy(ei[1])]
    ref = torch.relu(oracle_aggregate(ei, n, x, norm, "mean" if reduce else "add") + bias + res)
    assert_parity(out, ref, f"prescaled H={H}")
    # unweighted, non-hub rows: exactly the reference's edge-order sum
    plain = ops.aggregate_prescaled_impl(g.fwd, x.to(DEV), None, reduce, None, None, 0)
    refp = oracle_aggregate(ei, n, x, None, "mean" if reduce else "add")
    small = torch.from_numpy(np.bincount(ei[1], minlength=n) <= 64)
    assert_bitexact(plain.cpu()[small], refp[small], "plain non-hub rows")
    assert_parity(plain, refp, "plain all rows")
    assert_bitexact(ops.aggregate_prescaled_impl(g.fwd, xs, dis.to(DEV), reduce, bias.to(DEV), res.to(DEV), 1),
                    out, "deterministic")


def test_torch_library_ops_roundtrip():
    """the registered torch.ops.mgcn.* surface (list-of-tensors structure) matches the direct calls"""
    n, e = 500, 6000
    ei = torch.from_numpy(rand_graph(1, n, e)).to(DEV)
    csr = torch.ops.mgcn.csr_build(ei, n, 1, 0, 256)
    assert len(csr) == len(ops.CSR_FIELDS)
    x = torch.randn(n, 32, device=DEV)
    a = torch.ops.mgcn.spmm(csr, 256, x, False, None, None, None, 0, None, None, 0)
    b = ops.spmm_impl(GraphStructure(ei, n).fwd, x)
    assert_bitexact(a, b, "torch.ops.mgcn.spmm")
    c = torch.ops.mgcn.aggregate_prescaled(csr, 256, x, None, 0, None, None, 0)
    assert_bitexact(c, b, "torch.ops.mgcn.aggregate_prescaled")


def test_registered_ops_are_differentiable():
    """torch.ops.mgcn.propagate / linear / segment_reduce carry gradients through torch.library.register_autograd
    (SURVEY.md §8b): same values and gradients as the differentiable wrappers of meta_gcn_b200.functional"""
    from meta_gcn_b200 import functional as F
    from meta_gcn_b200.graph import GraphStructure
    n, e, h = 700, 5000, 32
    g = np.random.default_rng(5)
    ei = torch.from_numpy(np.stack([g.integers(0, n, e), g.integers(0, n, e)]).astype(np.int64)).to(DEV)
    gs = GraphStructure(ei, n, hub_threshold=64)
    gen = torch.Generator().manual_seed(1)
    x0 = torch.randn(n, h, generator=gen).to(DEV)
    w0 = (torch.randn(h, h, generator=gen) / 6).to(DEV)
    b0 = torch.randn(h, generator=gen).to(DEV)
    dis = (torch.rand(n, generator=gen) + 0.2).to(DEV)
    offsets = torch.tensor([0, 100, 350, n], dtype=torch.int32, device=DEV)
    wout = torch.randn(3, h, generator=gen).to(DEV)

    def run(use_ops):
        x = x0.clone().requires_grad_(True)
        w = w0.clone().requires_grad_(True)
        b = b0.clone().requires_grad_(True)
        if use_ops:
            y = torch.ops.mgcn.linear(x, w, False, b, None, 1)
            z = torch.ops.mgcn.propagate(gs.fwd.tensors(), gs.bwd.tensors(), 64, y, dis, dis, 1)
            p = torch.ops.mgcn.segment_reduce(z, offsets, 1)
        else:
            y = F.linear(x, w, b, act="relu")
            z = F.aggregate(y, gs, dis, dis, act="relu")
            p = F.segment_pool(z, offsets, "mean")
        (p * wout).sum().backward()
        return p.detach(), x.grad, w.grad, b.grad

    a, b = run(True), run(False)
    for u, v, name in zip(a, b, ("out", "dx", "dw", "db")):
        assert_bitexact(u, v, "torch.ops.mgcn autograd " + name)
    # the fused layer launch through the op table
    zn, hm = torch.ops.mgcn.gcn_layer_fwd_tc(gs.fwd.tensors(), 64, x0, w0, w0.t().contiguous(), b0, None, dis, dis, dis, 1)
    zn2, hm2 = ops.gcn_layer_fwd_tc_impl(gs.fwd, x0, w0, w0.t().contiguous(), b0, None, dis, dis, dis, 1)
    assert_bitexact(zn, zn2, "torch.ops.mgcn.gcn_layer_fwd_tc")
    assert_bitexact(hm, hm2, "torch.ops.mgcn.gcn_layer_fwd_tc mask")
    f = torch.ops.mgcn.edge_fingerprint(ei).tolist()
    assert (f[0] == f[1] and f[2] == f[3]) == ops.edge_symmetry_impl(ei)
