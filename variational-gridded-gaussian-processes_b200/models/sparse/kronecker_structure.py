"""2-D Kronecker-structured sparse models: drop-in names of src/models/sparse/kronecker_structure.py."""
from typing import Tuple

import torch

from ... import _lib
from ...basis import B0SplineBasis, B1SplineBasis
from .._gridded import GriddedVariationalGP, linspace_mesh


class KroneckerStructure(GriddedVariationalGP):
    """kronecker_structure.py:15-278 (base class name kept)."""


class _TwoDimMesh(KroneckerStructure):
    def __init__(self, X, y, nknots: int, dim1lims: Tuple[float, float], dim2lims: Tuple[float, float]):
        self.nknots = nknots
        self.dim1lims = dim1lims
        self.dim2lims = dim2lims
        mesh_1 = linspace_mesh(dim1lims, nknots)
        mesh_2 = linspace_mesh(dim2lims, nknots)
        super().__init__(X, y, [mesh_1, mesh_2])
        self.mesh_1, self.mesh_2 = mesh_1, mesh_2
        self.delta_1 = mesh_1[1] - mesh_1[0]
        self.delta_2 = mesh_2[1] - mesh_2[0]


class Matern12B1SplineASVGP(_TwoDimMesh):
    """kronecker_structure.py:524-662."""
    family = _lib.B1_ASVGP

    def __init__(self, X, y, nknots, dim1lims, dim2lims):
        super().__init__(X, y, nknots, dim1lims, dim2lims)
        self.delta = self.delta_1
        self.basis_1 = B1SplineBasis(self.mesh_1)
        self.basis_2 = B1SplineBasis(self.mesh_2)


class Matern12B0SplineGriddedGP(_TwoDimMesh):
    """kronecker_structure.py:671-849."""
    family = _lib.B0_GRIDDED

    def __init__(self, X, y, nknots, dim1lims, dim2lims):
        super().__init__(X, y, nknots, dim1lims, dim2lims)
        self.basis_1 = B0SplineBasis(self.mesh_1)
        self.basis_2 = B0SplineBasis(self.mesh_2)

    def q_v(self, optimal: bool = False):
        """q(v) = q(u): the inducing variables are the cell integrals (kronecker_structure.py:825-849 returns the
        closed-form optimum: `optimal=True`; default: the learned q(u))."""
        return self.q_u(optimal)
