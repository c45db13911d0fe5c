"""esc_gnn_b200 -- B200-native drop-in for ESC-GNN's two hot paths: the efficient structural-encoding transform and
the NestedGIN_eff train step.  See DESIGN.md / INTEGRATION.md.  No CPU fallback anywhere in this package."""
from .data import Data  # noqa: F401
from .batch import Batch  # noqa: F401
from .dataloader import DataLoader  # noqa: F401
from .transform import create_subgraphs, encode_batch, encode_batch_host  # noqa: F401

__all__ = ['Data', 'Batch', 'DataLoader', 'create_subgraphs', 'encode_batch', 'encode_batch_host']
