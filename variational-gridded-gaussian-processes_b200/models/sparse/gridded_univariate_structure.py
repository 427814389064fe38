"""1-D gridded models: drop-in names of src/models/sparse/gridded_univariate_structure.py."""
from typing import Tuple

import torch

from ... import _lib
from ...basis import B0SplineBasis, B1SplineBasis
from .._gridded import b0_cell_cov, linspace_mesh
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

    # ---- gridded part (gridded_univariate_structure.py:595-700), dense: 1-D problems are small --------------------
    def _Kvu(self) -> torch.Tensor:
        """L2 inner products of the B0 cell indicators with the B1 hats (:595-608): delta/2 on the two knots bounding a
        cell, delta on the n_b1_splines knots inside it."""
        half = (self.b1_basis_1.delta / 2.0).to(torch.float64)
        full = self.b1_basis_1.delta.to(torch.float64)
        nk, step = self.b1_basis_1.n_basis_functions, self.n_b1_splines + 1
        Kvu = torch.zeros(self.n_b0_splines, nk, dtype=torch.float64)
        for i in range(self.n_b0_splines):
            lo = step * (i + 1)                       # first knot of B0 cell i (one padding cell to the left)
            Kvu[i, lo] = half
            Kvu[i, lo + 1:lo + step] = full
            Kvu[i, lo + step] = half
        return Kvu.to(self.variational_mean.device)

    def _Kvv(self) -> torch.Tensor:
        """Toeplitz covariance of the cell integrals (:611-646)."""
        return b0_cell_cov(self.b0_delta_1, self.n_b0_splines, self.kernel.base_kernel.lengthscale,
                           self.kernel.outputscale).to(self.variational_mean.device)

    def _v_given(self, optimal: bool):
        Kuu, Kvu = self._Kuu(), self._Kvu()
        Lk = torch.linalg.cholesky(Kuu)
        B = torch.cholesky_solve(Kvu.T.contiguous(), Lk)                   # Kuu^-1 Kuv
        return (self.q_u_optimal() if optimal else self.q_u()), Kvu, B

    def p_v_u(self, optimal: bool = False):
        """p(v | u = E_q[u]) (:674-685)."""
        from ...params import DenseNormal
        q, Kvu, B = self._v_given(optimal)
        return DenseNormal(B.T @ q.mean.to(torch.float64), self._Kvv() - Kvu @ B)

    def q_v(self, optimal: bool = False):
        """q(v) (:687-700): mean = Kvu Kuu^-1 m, cov = Kvv - Kvu Kuu^-1 Kuv + Kvu Kuu^-1 S Kuu^-1 Kuv.  With the optimal
        q(u) (`optimal=True`) Kuu^-1 S* Kuu^-1 = Sigma^-1 and Kuu^-1 m* = Sigma^-1 Kuf y / noise: the reference's formula."""
        from ...params import DenseNormal
        q, Kvu, B = self._v_given(optimal)
        S = q.covariance_matrix.to(torch.float64)
        return DenseNormal(B.T @ q.mean.to(torch.float64), self._Kvv() - Kvu @ B + B.T @ S @ B)


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

    def q_v(self, optimal: bool = False):
        """q(v) = q(u): the inducing variables are the cell integrals (:820-844 returns the closed-form optimum)."""
        return self.q_u(optimal)
