"""Narrow dense transform, its gradients, relu backward and the segment reductions vs fp64 /
the oracle."""
import numpy as np
import pytest
import torch

from meta_gcn_b200 import functional as F_mgcn
from meta_gcn_b200 import ops
from oracle import port
from util import assert_bitexact, assert_parity

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("N", [1, 127, 128, 129, 1000, 5001])
@pytest.mark.parametrize("Hi,Ho", [(1, 32), (3, 64), (32, 32), (32, 2), (33, 47), (64, 64), (100, 256), (256, 256)])
def test_linear_forward(N, Hi, Ho):
    g = torch.Generator().manual_seed(N * 7 + Hi)
    x = torch.randn(N, Hi, generator=g)
    w = torch.randn(Hi, Ho, generator=g) / max(Hi, 1) ** 0.5
    b = torch.randn(Ho, generator=g)
    add = torch.randn(N, Ho, generator=g)
    ref64 = x.double() @ w.double()
    y = ops.linear_impl(x.to(DEV), w.to(DEV), False)
    assert_parity(y, ref64, "x@W")
    # nn.Linear layout + bias + add + relu
    y2 = ops.linear_impl(x.to(DEV), w.t().contiguous().to(DEV), True, b.to(DEV), add.to(DEV), 1)
    assert_parity(y2, torch.relu(ref64 + b.double() + add.double()), "relu(x@W^T + b + add)")
    # fp32 CPU matmul (the reference's sgemm) is within the same tolerance of our result
    assert_parity(y, x @ w, "vs torch CPU fp32")


@pytest.mark.parametrize("N", [1, 63, 64, 65, 4097, 20000])
@pytest.mark.parametrize("Hi,Ho", [(1, 32), (3, 64), (32, 32), (32, 2), (33, 47), (100, 256)])
@pytest.mark.parametrize("out_in", [False, True])
def test_linear_weight_gradient(N, Hi, Ho, out_in):
    g = torch.Generator().manual_seed(N + Hi * 3 + Ho)
    x = torch.randn(N, Hi, generator=g)
    gy = torch.randn(N, Ho, generator=g)
    dw, db = ops.linear_wgrad_impl(x.to(DEV), gy.to(DEV), out_in, True)
    ref = x.double().t() @ gy.double()
    assert_parity(dw, ref.t() if out_in else ref, "dW")
    assert_parity(db, gy.double().sum(0), "db")
    dw2, _ = ops.linear_wgrad_impl(x.to(DEV), gy.to(DEV), out_in, True)
    assert_bitexact(dw2, dw, "deterministic dW")


def test_linear_autograd_matches_torch():
    N, Hi, Ho = 3000, 32, 32
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, Hi, generator=g, requires_grad=True)
    w = (torch.randn(Ho, Hi, generator=g) / Hi ** 0.5).requires_grad_(True)
    b = torch.randn(Ho, generator=g, requires_grad=True)
    add = torch.randn(N, Ho, generator=g, requires_grad=True)
    wgt = torch.randn(N, Ho, generator=g)
    ref = torch.relu(torch.nn.functional.linear(x, w, b) + add)
    (ref * wgt).sum().backward()
    xd, wd, bd, ad = (t.detach().to(DEV).requires_grad_(True) for t in (x, w, b, add))
    out = F_mgcn.linear(xd, wd, bd, add=ad, act="relu", weight_layout="out_in")
    (out * wgt.to(DEV)).sum().backward()
    assert_parity(out, ref, "y")
    assert_parity(xd.grad, x.grad, "dx")
    assert_parity(wd.grad, w.grad, "dW")
    assert_parity(bd.grad, b.grad, "db")
    assert_parity(ad.grad, add.grad, "dadd")


def test_relu_backward_exact():
    y = torch.randn(1000, 33)
    y[::7] = 0
    g = torch.randn(1000, 33)
    out = ops.relu_backward_impl(g.to(DEV), y.to(DEV))
    assert_bitexact(out, torch.where(y > 0, g, torch.zeros_like(g)), "relu'")


@pytest.mark.parametrize("H", [1, 3, 32, 64, 100, 300])
@pytest.mark.parametrize("mode", ["add", "mean"])
def test_segment_reduce_small_graphs_bitexact(H, mode):
    sizes = [5, 0, 40, 1, 64, 17, 0, 33]
    n = sum(sizes)
    batch = torch.cat([torch.full((s,), i, dtype=torch.long) for i, s in enumerate(sizes)])
    x = torch.randn(n, H)
    out = F_mgcn.pool_by_batch(x.to(DEV), batch.to(DEV), len(sizes), mode)
    assert_bitexact(out, port.scatter_rows(mode, x, batch, len(sizes)), f"pool {mode} H={H}")
    # size inferred from batch (host sync, like PyG)
    out2 = F_mgcn.pool_by_batch(x.to(DEV), batch[: n - 33].to(DEV), None, mode)
    assert out2.shape[0] == 6


@pytest.mark.parametrize("H", [2, 32, 256])
def test_segment_reduce_large_graphs_and_backward(H):
    sizes = [5000, 3, 12000, 700]
    n = sum(sizes)
    batch = torch.cat([torch.full((s,), i, dtype=torch.long) for i, s in enumerate(sizes)])
    x = torch.randn(n, H, requires_grad=True)
    wgt = torch.randn(len(sizes), H)
    ref = port.scatter_rows("mean", x, batch, len(sizes))
    (ref * wgt).sum().backward()
    xd = x.detach().to(DEV).requires_grad_(True)
    out = F_mgcn.pool_by_batch(xd, batch.to(DEV), len(sizes), "mean")
    (out * wgt.to(DEV)).sum().backward()
    assert_parity(out, ref, "mean pool")
    assert_parity(xd.grad, x.grad, "mean pool backward")
    out_b = F_mgcn.pool_by_batch(xd, batch.to(DEV), len(sizes), "mean")
    assert_bitexact(out_b, out, "deterministic")


@pytest.mark.parametrize("sizes", [[150000], [40000, 90000], [30000, 2, 0, 70000]])
def test_segment_reduce_sliced_long_segments(sizes):
    """average segment > 4096 rows: several CTAs share a segment (slices added in order), still deterministic"""
    H = 32
    n = sum(sizes)
    batch = torch.cat([torch.full((s,), i, dtype=torch.long) for i, s in enumerate(sizes)])
    x = torch.randn(n, H)
    for mode in ("mean", "add"):
        ref = port.scatter_rows(mode, x.double(), batch, len(sizes))
        out = F_mgcn.pool_by_batch(x.to(DEV), batch.to(DEV), len(sizes), mode)
        assert_parity(out, ref, f"{mode} pool, sliced")
        assert_bitexact(F_mgcn.pool_by_batch(x.to(DEV), batch.to(DEV), len(sizes), mode), out, "deterministic")


@pytest.mark.parametrize("C", [2, 5, 16])
@pytest.mark.parametrize("reduction", ["mean", "sum"])
def test_cross_entropy_matches_torch(C, reduction):
    N = 70001
    g = torch.Generator().manual_seed(C)
    z = (torch.randn(N, C, generator=g) * 3).requires_grad_(True)
    y = torch.randint(0, C, (N,), generator=g)
    ref = torch.nn.CrossEntropyLoss(reduction=reduction)(z.double(), y)
    ref.backward()
    zd = z.detach().to(DEV).requires_grad_(True)
    loss = F_mgcn.cross_entropy(zd, y.to(DEV), reduction)
    (loss * 1.5).backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert_parity(zd.grad, 1.5 * z.grad, "dlogits")
    loss2 = F_mgcn.cross_entropy(zd, y.to(DEV), reduction)
    assert loss2.item() == loss.item()


def test_masked_scale_and_masked_operands():
    N, H = 5000, 32
    g = torch.Generator().manual_seed(1)
    a, m1, m2 = (torch.randn(N, H, generator=g) for _ in range(3))
    rs = torch.rand(N, generator=g)
    out = ops.masked_scale_impl(a.to(DEV), m1.to(DEV), m2.to(DEV), rs.to(DEV))
    ref = rs.view(-1, 1) * torch.where((m1 > 0) & (m2 > 0), a, torch.zeros(()))
    assert_bitexact(out, ref, "masked_scale")
    w = torch.randn(H, H, generator=g) / H ** 0.5
    y = ops.linear_impl(a.to(DEV), w.to(DEV), False, xmask=m1.to(DEV), row_scale=rs.to(DEV))
    assert_parity(y, rs.double().view(-1, 1) * ((a * (m1 > 0)).double() @ w.double()), "masked, row-scaled linear")
    dw, db = ops.linear_wgrad_impl(a.to(DEV), m2.to(DEV), True, True, gmask=m1.to(DEV))
    gm = (m2 * (m1 > 0)).double()
    assert_parity(dw, gm.t() @ a.double(), "masked wgrad")
    assert_parity(db, gm.sum(0), "masked bias grad")


@pytest.mark.parametrize("n,h,training", [(5146, 64, True), (37, 64, True), (1000, 48, False), (200001, 32, True), (1, 8, False)])
def test_batch_norm_matches_torch_fp64(n, h, training):
    """csrc/batchnorm.cu behind torch.nn.BatchNorm1d (kernel/gin.py:15): outputs, running statistics and every
    gradient against torch's own BatchNorm1d evaluated in fp64 on the CPU"""
    from meta_gcn_b200 import functional as F
    gen = torch.Generator().manual_seed(n + h)
    x = torch.randn(n, h, generator=gen) * 2.0 + 3.0        # mean^2 >> var: the one-pass variance would cancel
    wout = torch.randn(n, h, generator=gen)
    bn = torch.nn.BatchNorm1d(h)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(h, generator=gen) + 0.5)
        bn.bias.copy_(torch.randn(h, generator=gen))
        bn.running_mean.copy_(torch.randn(h, generator=gen))
        bn.running_var.copy_(torch.rand(h, generator=gen) + 0.5)
    import copy
    ref = copy.deepcopy(bn).double()
    bn = bn.to("cuda")
    bn.train(training)
    ref.train(training)
    if n == 1 and training:
        pytest.skip("torch refuses batch statistics of a single row")
    xr = x.double().requires_grad_(True)
    yr = ref(xr)
    (yr * wout.double()).sum().backward()
    xd = x.to("cuda").requires_grad_(True)
    y = F.batch_norm(xd, bn)
    (y * wout.to("cuda")).sum().backward()
    assert_parity(y, yr, "bn.y")
    assert_parity(xd.grad, xr.grad, "bn.dx")
    assert_parity(bn.weight.grad, ref.weight.grad, "bn.dgamma")
    assert_parity(bn.bias.grad, ref.bias.grad, "bn.dbeta")
    assert_parity(bn.running_mean, ref.running_mean, "bn.running_mean")
    assert_parity(bn.running_var, ref.running_var, "bn.running_var")
    assert int(bn.num_batches_tracked) == int(ref.num_batches_tracked)
