"""Cross-section workspaces on the B200 path: spin-1/2 elastic, quasi-elastic (p,n) DWBA."""
from . import elastic, quasielastic_pn

__all__ = ["elastic", "quasielastic_pn"]
