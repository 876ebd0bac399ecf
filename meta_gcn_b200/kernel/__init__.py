"""Host-side mirror of the reference's ``kernel/`` benchmark nets on the hot path
(kernel/gcn.py, gin.py, graph_sage.py) over the libmgcn operators."""
from .gcn import GCN, GCNWithJK  # noqa: F401
from .gin import GIN, GIN0, GIN0WithJK, GINWithJK  # noqa: F401
from .graph_sage import GraphSAGE, GraphSAGEWithJK  # noqa: F401
