"""ORACLE ONLY — Data / Batch.from_data_list contract (src/gcn_meta/data/dataloader.py:11,
data.py:6-26; kernel/train_eval.py:37-39): concatenate x / y along dim 0, edge_index along dim 1 with
a cumulative node offset, `batch` = graph id per node (sorted ascending)."""
import torch


class Data:
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, **kwargs):
        self.x, self.edge_index, self.edge_attr, self.y = x, edge_index, edge_attr, y
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None and not k.startswith("_")]

    def __getitem__(self, key):
        return getattr(self, key)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def __call__(self, *keys):
        for k in sorted(self.keys) if not keys else keys:
            yield k, self[k]

    @property
    def num_nodes(self):
        if getattr(self, "x", None) is not None:
            return self.x.size(0)
        return int(self.edge_index.max()) + 1

    def to(self, device):
        for k in self.keys:
            v = self[k]
            if torch.is_tensor(v):
                self[k] = v.to(device)
        return self


class Batch(Data):
    @staticmethod
    def from_data_list(data_list):
        keys = data_list[0].keys
        batch = Batch()
        slices = {k: [0] for k in keys}
        cols = {k: [] for k in keys}
        ids = []
        offset = 0
        for i, d in enumerate(data_list):
            n = d.num_nodes
            for k in keys:
                v = d[k]
                if k == "edge_index":
                    v = v + offset
                cols[k].append(v)
                dim = -1 if k == "edge_index" else 0
                slices[k].append(slices[k][-1] + (v.size(dim) if torch.is_tensor(v) and v.dim() > 0 else 1))
            ids.append(torch.full((n,), i, dtype=torch.long))
            offset += n
        for k in keys:
            v0 = cols[k][0]
            if torch.is_tensor(v0):
                batch[k] = torch.cat(cols[k], dim=-1 if k == "edge_index" else 0) if v0.dim() > 0 \
                    else torch.stack(cols[k])
            else:
                batch[k] = torch.tensor(cols[k])
        batch.batch = torch.cat(ids)
        batch.__slices__ = slices
        return batch

    @property
    def num_graphs(self):
        return int(self.batch.max()) + 1
