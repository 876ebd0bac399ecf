"""INTEGRATION.md path 1: the reference's OWN files import over ``meta_gcn_b200.compat.install()`` — the import
surface of kernel/gcn.py:4, gin.py:4, graph_sage.py:4, src/gcn_meta/models/{common,gcn_base_models,gcn_multi_kernel,
gcn_model}.py.  /root/reference exists only in the build container, so these tests skip on the GPU box; a subprocess
keeps the sys.modules aliases out of the other tests."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this box")

_IMPORTS = r"""
import importlib.util, os, sys
sys.path.insert(0, {root!r})
from meta_gcn_b200 import compat
compat.install(force=True)
sys.path.insert(1, os.path.join({ref!r}, "src"))
def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
import torch
gcn = load("ref_kernel_gcn", os.path.join({ref!r}, "kernel", "gcn.py"))
gin = load("ref_kernel_gin", os.path.join({ref!r}, "kernel", "gin.py"))
sage = load("ref_kernel_sage", os.path.join({ref!r}, "kernel", "graph_sage.py"))
from gcn_meta.models.gcn_model import GCNModel            # the reference's class, over compat's torch_scatter / inits
from gcn_meta.models.common import scatter_
class DS:
    num_features, num_classes = 3, 2
nets = [gcn.GCN(DS(), 3, 64), gcn.GCNWithJK(DS(), 3, 64), gin.GIN0(DS(), 3, 64), gin.GIN(DS(), 3, 64),
        sage.GraphSAGE(DS(), 3, 64)]
import meta_gcn_b200.compat.torch_geometric.nn as cnn
assert type(nets[0].conv1) is cnn.GCNConv and type(nets[4].conv1) is cnn.SAGEConv and type(nets[2].conv1) is cnn.GINConv
m = GCNModel(in_channels=1, enc_sizes=[32] * 3, num_classes=2, residual_hop=1, dropout=0.0, final_type="proj",
             deg_norm="sm", aggr="add", bias=False)
assert sum(p.numel() for p in m.parameters()) == 32 + 2 * 1024 + 64 + 2 * 1056 + 66
print("IMPORT-OK")
{run}
"""

_RUN_CUDA = r"""
from meta_gcn_b200 import data as D
dev = "cuda"
tb = D.synth_tu_batch(seed=0, num_graphs=16).to(dev)
for net in nets:
    net = net.to(dev).train()
    out = net(tb)
    torch.nn.functional.nll_loss(out, tb.y.view(-1).long()).backward()
    assert torch.isfinite(out).all() and out.shape == (16, 2)
g = D.synth_botnet_graph(seed=1, num_nodes=3000, edge_entries=30000, evil=200)
x = torch.from_numpy(g["x"]).to(dev); ei = torch.from_numpy(g["edge_index"]).to(dev)
m = m.to(dev)
out = m(x[:, 0:1].contiguous(), ei, deg_K=x[:, 1].contiguous())       # reference GCNModel.forward on libmgcn scatter ops
torch.nn.CrossEntropyLoss()(out, torch.from_numpy(g["y"]).long().to(dev)).backward()
from oracle import port
ref = port.OracleGCNModel(in_channels=1, enc_sizes=[32] * 3, num_classes=2, residual_hop=1, dropout=0.0,
                          final_type="proj", deg_norm="sm", aggr="add", bias=False)
ref.load_state_dict(m.state_dict())
want = ref(x.cpu()[:, 0:1], ei.cpu(), x.cpu()[:, 1])
err = (out.detach().cpu() - want.detach()).abs().max().item() / want.abs().max().item()
assert err < 1e-5, err
print("RUN-OK", err)
"""


def _run(code):
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)


@needs_ref
def test_reference_files_import_over_the_compat_namespaces():
    res = _run(_IMPORTS.format(root=ROOT, ref=REF, run=""))
    assert res.returncode == 0 and "IMPORT-OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]


@needs_ref
@pytest.mark.gpu
def test_reference_files_run_on_libmgcn_through_compat():
    res = _run(_IMPORTS.format(root=ROOT, ref=REF, run=_RUN_CUDA))
    assert res.returncode == 0 and "RUN-OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
