"""Cross-section workspaces (spin-1/2 elastic) on the B200 path."""
from . import elastic

__all__ = ["elastic"]
