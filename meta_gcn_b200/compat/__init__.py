"""Drop-in namespaces for exactly the third-party symbols the reference's hot-path files import
(SURVEY.md §8b): ``torch_scatter.{scatter_add,scatter_mean}`` and
``torch_geometric.nn.{GCNConv,SAGEConv,GINConv,global_mean_pool,global_add_pool,JumpingKnowledge}``,
``torch_geometric.nn.inits.{glorot,zeros}``, backed by libmgcn (CUDA only).

``install()`` registers them in ``sys.modules`` under the third-party names, so the reference's own
files (kernel/gcn.py, gin.py, graph_sage.py, src/gcn_meta/models/*.py) import unmodified and run on
the B200 kernels — see INTEGRATION.md."""
import sys


def install(force=False):
    from . import torch_geometric as tg
    from . import torch_scatter as ts
    from .torch_geometric import data as tg_data
    from .torch_geometric import nn as tg_nn
    from .torch_geometric import utils as tg_utils
    from .torch_geometric.nn import inits as tg_inits
    mods = {
        "torch_scatter": ts,
        "torch_geometric": tg,
        "torch_geometric.nn": tg_nn,
        "torch_geometric.nn.inits": tg_inits,
        "torch_geometric.utils": tg_utils,
        "torch_geometric.data": tg_data,
    }
    for name, mod in mods.items():
        if force or name not in sys.modules:
            sys.modules[name] = mod
    return mods
