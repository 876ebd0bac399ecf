"""Wide dense transform on tcgen05 / TMEM (csrc/linear_wide.cu, 3xTF32) against an fp64 product on the same
seeded inputs: SAGE / GCN widths, ragged row counts (N not a multiple of the 128-row tile), K not a multiple of
the 32-column chunk (ogbn-products' 100 input features), both weight layouts, fused bias / add / ReLU / row scale."""
import pytest
import torch

from meta_gcn_b200 import ops
from util import assert_parity

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("n,hi,ho", [(1000, 256, 256), (300, 100, 256), (129, 64, 64), (5000, 256, 128),
                                     (128, 32, 192), (1, 256, 256), (40000, 256, 256)])
@pytest.mark.parametrize("out_in", [False, True])
def test_linear_wide_matches_fp64(n, hi, ho, out_in):
    g = torch.Generator().manual_seed(n + hi + ho)
    x = torch.randn(n, hi, generator=g)
    w = torch.randn((ho, hi) if out_in else (hi, ho), generator=g) / hi ** 0.5
    b = torch.randn(ho, generator=g)
    add = torch.randn(n, ho, generator=g)
    rs = torch.rand(n, generator=g) + 0.5
    d = lambda t: t.to(DEV)
    wm = (w.t() if out_in else w).double()
    y = ops.linear_impl(d(x), d(w), out_in)
    assert_parity(y, x.double() @ wm, "x W")
    y = ops.linear_impl(d(x), d(w), out_in, d(b), d(add), 1, None, d(rs))
    ref = rs.double().view(-1, 1) * torch.relu(x.double() @ wm + b.double() + add.double())
    assert_parity(y, ref, "rs * relu(x W + b + add)")


def test_linear_wide_is_deterministic():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3000, 256, generator=g).to(DEV)
    w = torch.randn(256, 256, generator=g).to(DEV)
    a = ops.linear_impl(x, w, False)
    b = ops.linear_impl(x, w, False)
    assert torch.equal(a, b)


@pytest.mark.parametrize("n,hi,ho,masked,out_in", [
    (40000, 256, 256, False, False),
    (65537, 100, 256, True, True),       # SAGE first layer at C4: 100 -> 256, ragged row count, masked gradient
    (33000, 64, 64, False, True),
    (70011, 128, 192, True, False),
    (150000, 256, 128, False, False),    # several accumulator chains per CTA
])
def test_wide_weight_gradient_tcgen05(n, hi, ho, masked, out_in):
    """dW = x^T (g * (mask > 0)), db = colsum on tcgen05 (csrc/wgrad_wide.cu: both operands MN-major, 3xTF32,
    chains cut every 16 chunks) against fp64 — autograd of SAGEConv / GCNConv matmul(x, weight) + bias
    (kernel/graph_sage.py:10,13; kernel/gcn.py:10,13)"""
    gen = torch.Generator().manual_seed(n + hi)
    x = torch.randn(n, hi, generator=gen)
    g = torch.randn(n, ho, generator=gen)
    m = torch.randn(n, ho, generator=gen) if masked else None
    dw, db = ops.linear_wgrad_impl(x.to(DEV), g.to(DEV), out_in, True, gmask=m.to(DEV) if masked else None)
    gm = g.double() * (m > 0) if masked else g.double()
    ref = x.double().t() @ gm
    assert_parity(dw, ref.t() if out_in else ref, "dW")
    assert_parity(db, gm.sum(0), "db")
    dw2, db2 = ops.linear_wgrad_impl(x.to(DEV), g.to(DEV), out_in, True, gmask=m.to(DEV) if masked else None)
    assert torch.equal(dw, dw2) and torch.equal(db, db2)      # deterministic: no atomics
