"""1-D sparse models: drop-in names of src/models/sparse/univariate_structure.py."""
from typing import Tuple

from ... import _lib
from ...basis import B0SplineBasis, B1SplineBasis
from .._gridded import GriddedVariationalGP, linspace_mesh


class SparseGP(GriddedVariationalGP):
    """univariate_structure.py:15-263 (base class name kept)."""


class Matern12B1SplineASVGP(SparseGP):
    """univariate_structure.py:531-658."""
    family = _lib.B1_ASVGP

    def __init__(self, X, y, nknots: int, dim1lims: Tuple[float, float]):
        self.nknots = nknots
        self.alim, self.blim = dim1lims[0], dim1lims[1]
        mesh = linspace_mesh(dim1lims, nknots)
        super().__init__(X, y, [mesh])
        self.mesh = mesh
        self.delta = mesh[1] - mesh[0]
        self.basis = B1SplineBasis(mesh)


class Matern12B0SplineGriddedGP(SparseGP):
    """univariate_structure.py:668-825."""
    family = _lib.B0_GRIDDED

    def __init__(self, X, y, nknots: int, dim1lims: Tuple[float, float]):
        self.nknots = nknots
        self.alim, self.blim = dim1lims[0], dim1lims[1]
        mesh = linspace_mesh(dim1lims, nknots)
        super().__init__(X, y, [mesh])
        self.mesh = mesh
        self.delta = mesh[1] - mesh[0]
        self.basis = B0SplineBasis(mesh)
