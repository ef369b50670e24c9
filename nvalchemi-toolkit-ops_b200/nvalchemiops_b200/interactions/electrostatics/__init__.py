from .coulomb import coulomb_energy, coulomb_energy_forces, coulomb_forces, fused_coulomb_energy_forces  # noqa: F401

__all__ = ["coulomb_energy", "coulomb_forces", "coulomb_energy_forces", "fused_coulomb_energy_forces"]
