from .attention_ncf import AttentionNCF  # noqa: F401
from .basic_ncf import BasicNCF  # noqa: F401
from .gnn_ncf import GraphNCF  # noqa: F401
