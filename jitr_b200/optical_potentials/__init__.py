"""Built-in optical-potential parameterisations whose forms are evaluated on the device."""
from . import chuq, dom, kduq, nonlocal_forms, wlh
from .omp import LocalOpticalPotential, SingleChannelOpticalModel, ShapeModel

__all__ = ["LocalOpticalPotential", "ShapeModel", "SingleChannelOpticalModel", "chuq", "dom", "kduq", "nonlocal_forms", "wlh"]
