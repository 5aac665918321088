"""Channel-system objects and reaction descriptors consumed by the R-matrix path."""
from .reaction import ElasticReaction, Particle, Reaction
from .system import Asymptotics, Channels, ProjectileTargetSystem, scalar_couplings, spin_half_orbit_coupling
from .wavefunction import Wavefunctions

__all__ = [
    "Asymptotics", "Channels", "ElasticReaction", "Particle", "ProjectileTargetSystem", "Reaction",
    "Wavefunctions", "scalar_couplings", "spin_half_orbit_coupling",
]
