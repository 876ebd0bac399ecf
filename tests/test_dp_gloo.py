"""Graph-sharded data parallelism on CPU (gloo, world_size 2): the flat-buffer all-reduce with
global-node-count normalisation reproduces the single-process batched step.  The compute inside each
rank is the oracle model (the CUDA ops need a GPU); what is under test is meta_gcn_b200.dist."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from meta_gcn_b200 import data as D
from meta_gcn_b200 import dist as mdist
from oracle import port

CFG = dict(in_channels=1, enc_sizes=[8, 8, 8], num_classes=2, residual_hop=1, dropout=0.0,
           final_type="proj", deg_norm="sm", aggr="add", bias=False)


def _graphs():
    return [D.synth_botnet_graph(seed=s, num_nodes=300 + 40 * s, edge_entries=3000 + 500 * s, evil=30)
            for s in range(5)]


def _step(model, batch, reducer):
    reducer.zero()
    out = model(batch.x[:, 0:1].contiguous(), batch.edge_index, batch.x[:, 1].contiguous())
    loss_sum = torch.nn.functional.cross_entropy(out, batch.y.long(), reduction="sum")
    loss_sum.backward()
    return reducer.reduce_mean(loss_sum, batch.num_nodes)


def _worker(rank, world, port_no, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    r, _, w = mdist.init_from_env("gloo")
    graphs = _graphs()
    a, b = mdist.shard_range(len(graphs), w, r)
    local = D.GraphBatch.from_data_list(graphs[a:b])
    torch.manual_seed(0)
    model = port.OracleGCNModel(**CFG)
    reducer = mdist.FlatGradientReducer(model.parameters())
    mean_loss, total = _step(model, local, reducer)
    if r == 0:
        ret["loss"] = float(mean_loss)
        ret["total"] = float(total)
        ret["grads"] = reducer.grads.clone().numpy()
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_step_equals_batched_step():
    torch.set_num_threads(1)
    graphs = _graphs()
    full = D.GraphBatch.from_data_list(graphs)
    torch.manual_seed(0)
    model = port.OracleGCNModel(**CFG)
    out = model(full.x[:, 0:1].contiguous(), full.edge_index, full.x[:, 1].contiguous())
    loss = torch.nn.CrossEntropyLoss()(out, full.y.long())          # train_botnet.py:225,287
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).numpy()

    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret["total"] == full.num_nodes
    assert abs(ret["loss"] - loss.item()) < 1e-5
    got = ret["grads"]
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 1e-5 * scale, np.abs(got - ref).max() / scale


def test_single_process_reducer_is_identity_mean():
    torch.manual_seed(0)
    model = port.OracleGCNModel(**CFG)
    red = mdist.FlatGradientReducer(model.parameters())
    g = D.GraphBatch.from_data_list(_graphs()[:1])
    mean_loss, total = _step(model, g, red)
    assert float(total) == g.num_nodes
    # grads are views into the flat buffer
    off = 0
    for p in model.parameters():
        assert p.grad.data_ptr() == red.flat[off:off + p.numel()].data_ptr()
        off += p.numel()
