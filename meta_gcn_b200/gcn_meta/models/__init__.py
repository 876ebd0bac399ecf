from .common import activation, scatter_, softmax  # noqa: F401
from .gcn_base_models import NodeModelAdditive, NodeModelBase  # noqa: F401
from .gcn_model import GCNLayer, GCNModel  # noqa: F401
from .gcn_multi_kernel import GCNMultiKernel  # noqa: F401
from .graph_attention import NodeModelAttention  # noqa: F401
