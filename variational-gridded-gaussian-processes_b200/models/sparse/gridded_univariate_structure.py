"""1-D gridded models: drop-in names of src/models/sparse/gridded_univariate_structure.py."""
from typing import Tuple

import torch

from ... import _lib
from ...basis import B0SplineBasis, B1SplineBasis
from .._gridded import linspace_mesh
from .univariate_structure import SparseGP


class GriddedMatern12ASVGP(SparseGP):
    """gridded_univariate_structure.py:497-700: B1 sub-mesh (n_b1_splines hats per B0 cell) on the B0 mesh padded
    by one cell each side."""
    family = _lib.B1_ASVGP

    def __init__(self, X, y, n_b0_splines: int, n_b1_splines: int, dimlims: Tuple[float, float]):
        self.dimlims = dimlims
        self.n_b0_splines = n_b0_splines
        self.n_b1_splines = n_b1_splines
        self.padding = 1
        b0 = torch.linspace(dimlims[0], dimlims[1], n_b0_splines + 1)
        d = b0[1] - b0[0]
        padded = torch.cat(((b0[0] - d).unsqueeze(-1), b0, (b0[-1] + d).unsqueeze(-1)))
        b1 = torch.stack([torch.linspace(padded[i], padded[i + 1], n_b1_splines + 2)[:-1]
                          for i in range(n_b0_splines + 2 * self.padding)], dim=0).flatten()
        b1 = torch.cat((b1, padded[-1].view(1)))
        super().__init__(X, y, [b1])
        self.b0_mesh_1, self.b0_delta_1 = b0, d
        self.b0_mesh_padded_1 = padded
        self.b1_mesh_1 = b1
        self.b0_basis_1 = B0SplineBasis(b0)
        self.b1_basis_1 = B1SplineBasis(b1)


class Matern12GriddedGP(SparseGP):
    """gridded_univariate_structure.py:709-844."""
    family = _lib.B0_GRIDDED

    def __init__(self, X, y, n_b0_splines: int, gridlims: Tuple[float, float]):
        self.n_b0_splines = n_b0_splines
        self.gridlims = gridlims
        mesh = torch.linspace(gridlims[0], gridlims[1], n_b0_splines + 1)
        super().__init__(X, y, [mesh])
        self.b0_mesh_1 = mesh
        self.b0_basis = B0SplineBasis(mesh)

    def q_v(self):
        return self.q_u()
