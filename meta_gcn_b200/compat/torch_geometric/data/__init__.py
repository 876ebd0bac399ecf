"""``torch_geometric.data`` stand-ins: Batch.from_data_list contract (dataloader.py:11)."""
from ....data import GraphBatch


class Data(dict):
    """attribute bag: Data(x=..., edge_index=..., y=...)"""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.__dict__ = self


class Batch(GraphBatch):
    pass
