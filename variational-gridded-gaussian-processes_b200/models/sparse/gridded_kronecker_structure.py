"""2-D gridded (grid-product) models: drop-in names of src/models/sparse/gridded_kronecker_structure.py."""
from typing import Tuple

import torch

from ... import _lib
from ...basis import B0SplineBasis, B1SplineBasis
from .._gridded import GriddedVariationalGP, linspace_mesh, padded_b0_mesh
from .kronecker_structure import KroneckerStructure


class GriddedMatern12ASVGP(KroneckerStructure):
    """gridded_kronecker_structure.py:685-969: B1 features on the padded B0 mesh."""
    family = _lib.B1_ASVGP

    def __init__(self, X, y, n_b0_splines: int, padding_factor: int,
                 dim1_grid_lims: Tuple[float, float], dim2_grid_lims: Tuple[float, float]):
        self.dim1_grid_lims = dim1_grid_lims
        self.dim2_grid_lims = dim2_grid_lims
        self.n_b0_splines = n_b0_splines
        self.padding_factor = padding_factor
        b0_1, d1, pad_1 = padded_b0_mesh(dim1_grid_lims, n_b0_splines, padding_factor)
        b0_2, d2, pad_2 = padded_b0_mesh(dim2_grid_lims, n_b0_splines, padding_factor)
        super().__init__(X, y, [pad_1, pad_2])
        self.b0_mesh_1, self.b0_mesh_2 = b0_1, b0_2
        self.b0_delta_1, self.b0_delta_2 = d1, d2
        self.b0_mesh_padded_1, self.b0_mesh_padded_2 = pad_1, pad_2
        self.b0_basis_1, self.b0_basis_2 = B0SplineBasis(b0_1), B0SplineBasis(b0_2)
        self.b1_basis_1, self.b1_basis_2 = B1SplineBasis(pad_1), B1SplineBasis(pad_2)


class Matern12GriddedGP(KroneckerStructure):
    """gridded_kronecker_structure.py:1255-1433: the inducing variables are the B0 cell integrals."""
    family = _lib.B0_GRIDDED

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
        self.basis_1, self.basis_2 = B0SplineBasis(mesh_1), B0SplineBasis(mesh_2)

    def q_v(self):
        """For this model the inducing variables are the gridded cell integrals: q(v) is q(u) (:1409-1433)."""
        return self.q_u()
