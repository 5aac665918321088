"""Small host-side helpers of the path (kinematics, constants, block slicing)."""
from .. import constants
from .kinematics import ChannelKinematics, classical_kinematics, classical_kinematics_cm, semi_relativistic_kinematics
from .misc import block, delta, interaction_range, smatrix, suggested_basis_size, suggested_dimensionless_channel_radius

__all__ = [
    "constants", "ChannelKinematics", "classical_kinematics", "classical_kinematics_cm",
    "semi_relativistic_kinematics", "block", "delta", "interaction_range", "smatrix",
    "suggested_basis_size", "suggested_dimensionless_channel_radius",
]
