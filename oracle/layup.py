"""Oracle: composite layup through-thickness integration.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
Follows src/CompositeLayupModule.jl (reference v3.6.4).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import fe_external as fx
from .shells import plane_stress_T, shell_material_stiffness, transverse_shear_T


@dataclass
class Ply:
    """src/CompositeLayupModule.jl:53-97 -- `D6` is the ply material's 3-D moduli."""

    name: str
    D6: np.ndarray
    thickness: float
    angle: float  # degrees
    rho: float = 0.0
    Dps: np.ndarray = field(init=False)
    Dts: np.ndarray = field(init=False)

    def __post_init__(self):
        self.Dps, self.Dts = shell_material_stiffness(self.D6)


def lamina_moduli(E1, E2, nu12, G12, G13, G23):
    """`lamina_material(E1,E2,nu12,G12,G13,G23)`: E3=E2, nu13=nu12, nu23=0.
    src/CompositeLayupModule.jl:105-122."""
    return fx.moduli_ortho(E1, E2, E2, nu12, nu12, 0.0, G12, G13, G23)


def cartesian_csys(axes):
    """`cartesian_csys(axes)` -> constant 3x3 csmat.  src/CompositeLayupModule.jl:26-44."""
    M = np.zeros((3, 3))
    for j in range(3):
        M[abs(axes[j]) - 1, j] = np.sign(axes[j]) * 1.0
    return M


@dataclass
class CompositeLayup:
    """src/CompositeLayupModule.jl:185-225."""

    name: str
    plies: list
    offset: float = 0.0
    vinson_sierakowski: bool = True
    transverse_shear_constant: float = 0.0

    @property
    def thickness(self):
        """:232-234"""
        return sum(p.thickness for p in self.plies)

    def laminate_stiffnesses(self):
        """A, B, D.  :246-271"""
        A = np.zeros((3, 3))
        B = np.zeros((3, 3))
        D = np.zeros((3, 3))
        zs = -self.thickness / 2 - self.offset
        for p in self.plies:
            ze = zs + p.thickness
            T = plane_stress_T(p.angle / 180 * np.pi)
            Dps = T @ (p.Dps @ T.T)  # TransformerQEQt
            A += (ze - zs) * Dps
            B += (ze**2 - zs**2) / 2 * Dps
            D += (ze**3 - zs**3) / 3 * Dps
            zs += p.thickness
        return A, B, D

    def laminate_transverse_stiffness(self):
        """H.  :280-308"""
        H = np.zeros((2, 2))
        lt = self.thickness
        zs = -lt / 2 - self.offset
        for p in self.plies:
            ze = zs + p.thickness
            a = p.angle / 180 * np.pi
            T = transverse_shear_T(np.cos(a), np.sin(a))
            Dts = T.T @ (p.Dts @ T)  # TransformerQtEQ
            if self.vinson_sierakowski:
                H += 5 / 4 * (ze - zs - 4 / 3 * (ze**3 - zs**3) / lt**2) * Dts
            else:
                H += self.transverse_shear_constant * (ze - zs) * Dts
            zs += p.thickness
        return H

    def laminate_inertia(self):
        """(mass density, moment-of-inertia density).  :317-330"""
        zs = -self.thickness / 2 - self.offset
        md = 0.0
        mi = 0.0
        for p in self.plies:
            ze = zs + p.thickness
            md += (ze - zs) * p.rho
            mi += (ze**3 - zs**3) * p.rho / 3
            zs += p.thickness
        return md, mi
