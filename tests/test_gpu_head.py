"""Fused output layer + cross entropy + binary counters (csrc/head.cu) against the separate steps the
reference takes (gcn_model.py:108 nn.Linear, train_botnet.py:287 CrossEntropyLoss, optim/metrics.py:8-24) in fp64."""
import numpy as np
import pytest
import torch

from meta_gcn_b200 import functional as F
from meta_gcn_b200 import ops
from util import assert_bitexact, assert_parity

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("n,h,c,reduction", [(1000, 32, 2, "mean"), (100003, 32, 2, "sum"), (517, 64, 5, "mean"),
                                             (33, 16, 3, "sum"), (4097, 128, 8, "mean"), (1, 32, 2, "mean")])
def test_head_cross_entropy_matches_separate_steps(n, h, c, reduction):
    gen = torch.Generator().manual_seed(n + h + c)
    x = torch.randn(n, h, generator=gen)
    lin = torch.nn.Linear(h, c)
    y = torch.randint(0, c, (n,), generator=gen)
    upstream = 1.7
    xr = x.double().requires_grad_(True)
    ref = torch.nn.Linear(h, c).double()
    ref.load_state_dict({k: v.double() for k, v in lin.state_dict().items()})
    logits_r = ref(xr)
    loss_r = torch.nn.CrossEntropyLoss(reduction=reduction)(logits_r, y)
    (loss_r * upstream).backward()
    lin = lin.to(DEV)
    xd = x.to(DEV).requires_grad_(True)
    loss, logits, counts = F.head_cross_entropy(xd, lin, y.to(DEV), reduction, confusion=True)
    (loss * upstream).backward()
    assert_parity(logits, logits_r, "logits")
    assert abs(loss.item() - loss_r.item()) <= 1e-5 * max(1.0, abs(loss_r.item()))
    assert_parity(xd.grad, xr.grad, "dx")
    assert_parity(lin.weight.grad, ref.weight.grad, "dF")
    assert_parity(lin.bias.grad, ref.bias.grad, "df")
    want = ops.binary_confusion_impl(y.to(DEV), logits=logits)
    assert_bitexact(counts, want, "counters")
    # deterministic: a second run gives the same bits
    xd2 = x.to(DEV).requires_grad_(True)
    lin.zero_grad()
    loss2, logits2, _ = F.head_cross_entropy(xd2, lin, y.to(DEV), reduction, confusion=True)
    (loss2 * upstream).backward()
    assert_bitexact(loss2, loss, "loss run-to-run")
    assert_bitexact(xd2.grad, xd.grad, "dx run-to-run")


def test_forward_loss_equals_forward_plus_loss():
    """GCNModel.forward_loss == GCNModel.forward + CrossEntropyLoss (train_botnet.py:286-287), values and gradients"""
    from meta_gcn_b200 import data as D
    from meta_gcn_b200.gcn_meta.models import GCNModel
    cfg = dict(in_channels=1, enc_sizes=[32] * 4, num_classes=2, residual_hop=1, dropout=0.0, final_type="proj",
               deg_norm="sm", aggr="add", bias=False)
    g = D.synth_botnet_graph(seed=2, num_nodes=5000, edge_entries=50000, evil=300)
    b = D.GraphBatch.from_data_list([g]).to(DEV)
    x0, deg, y = b.x[:, 0:1].contiguous(), b.x[:, 1].contiguous(), b.y.long()
    torch.manual_seed(0)
    model = GCNModel(**cfg).to(DEV)
    out = model(x0, b.edge_index, deg_K=deg)
    loss_a = torch.nn.CrossEntropyLoss()(out, y)
    loss_a.backward()
    grads_a = [p.grad.clone() for p in model.parameters()]
    model.zero_grad()
    loss_b, logits, counts = model.forward_loss(x0, b.edge_index, y, deg_K=deg, confusion=True)
    loss_b.backward()
    assert_parity(logits, out, "logits")
    assert abs(loss_a.item() - loss_b.item()) <= 1e-6
    for ga, p in zip(grads_a, model.parameters()):
        assert_parity(p.grad, ga, "grad")
    assert int(counts[4]) == int((out.argmax(1) == y).sum())
