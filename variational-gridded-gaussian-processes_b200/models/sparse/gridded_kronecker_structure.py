"""2-D gridded (grid-product) models: drop-in names of src/models/sparse/gridded_kronecker_structure.py."""
from typing import Tuple

import torch

from ... import _lib
from ...basis import B0SplineBasis, B1SplineBasis
from .._gridded import GriddedVariationalGP, b0_cell_cov, linspace_mesh, padded_b0_mesh
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


    # ---- dense pieces of the gridded part, as the reference builds them (small problems) ------------------------
    def _Kvu_along_dim(self, dim: int) -> torch.Tensor:
        """Rows [delta, delta] on the two knots of each B0 cell (gridded_kronecker_structure.py:831-839; the reference's
        values are kept)."""
        basis = (self.b1_basis_1, self.b1_basis_2)[dim]
        delta = basis.delta.to(torch.float64)
        nb, pad, nk = self.n_b0_splines, self.padding_factor, basis.n_basis_functions
        Kvu = torch.zeros(nb, nk, dtype=torch.float64)
        i = torch.arange(nb)
        Kvu[i, pad + i] = delta
        Kvu[i, pad + i + 1] = delta
        return Kvu

    def _Kvu(self) -> torch.Tensor:
        """torch.kron(Kvu_1, Kvu_2) (:841-845)."""
        dev = self.variational_mean.device
        return torch.kron(self._Kvu_along_dim(0), self._Kvu_along_dim(1)).to(dev)

    def _Kvv_along_dim(self, dim: int) -> torch.Tensor:
        """Toeplitz covariance of the cell integrals along one dimension (:847-885)."""
        k = self._kernels[dim]
        return b0_cell_cov((self.b0_delta_1, self.b0_delta_2)[dim], self.n_b0_splines, k.base_kernel.lengthscale, k.outputscale)

    def _Kvv(self) -> torch.Tensor:
        """torch.kron(Kvv_1, Kvv_2) (:887-901)."""
        dev = self.variational_mean.device
        return torch.kron(self._Kvv_along_dim(0), self._Kvv_along_dim(1)).to(dev)

    def p_v_u(self, optimal: bool = False):
        """p(v | u = E_q[u]) (:918-928): mean = Kvu Kuu^-1 m, cov = Kvv - Kvu Kuu^-1 Kuv; dense."""
        from ...params import DenseNormal
        self._dense_guard("p_v_u()")
        Kuu, Kvu = self._Kuu(), self._Kvu()
        Lk = torch.linalg.cholesky(Kuu)
        B = torch.cholesky_solve(Kvu.T.contiguous(), Lk)                   # Kuu^-1 Kuv
        mean_u = (self.q_u_optimal() if optimal else self.q_u()).mean.to(torch.float64)
        return DenseNormal(B.T @ mean_u, self._Kvv() - Kvu @ B)

    def q_v_dense(self, optimal: bool = False):
        """q(v) with its full covariance (:930-947): mean = Kvu Kuu^-1 m, cov = Kvv - Kvu Kuu^-1 Kuv + Kvu Kuu^-1 S Kuu^-1 Kuv
        (the reference writes S^-1 in the last term, its 1-D twin Sigma^-1 = Kuu^-1 S* Kuu^-1: the latter is implemented)."""
        from ...params import DenseNormal
        self._dense_guard("q_v_dense()")
        q = self.q_u_optimal() if optimal else self.q_u()
        Kuu, Kvu = self._Kuu(), self._Kvu()
        Lk = torch.linalg.cholesky(Kuu)
        B = torch.cholesky_solve(Kvu.T.contiguous(), Lk)
        S = q.covariance_matrix.to(torch.float64)
        return DenseNormal(B.T @ q.mean.to(torch.float64), self._Kvv() - Kvu @ B + B.T @ S @ B)

    def q_v(self):
        """q(v) for the B0 cell integrals v_c = int_cell f (gridded_kronecker_structure.py:831-947), marginals only:
            mean = Kvu Kuu^-1 m = Kvu alpha,      Kvu = kron_d Kvu_d, rows of Kvu_d = [delta, delta] on the two knots of a cell
            var  = diag(Kvv) - diag(Kvu Kuu^-1 Kuv) + diag(Kvu Kuu^-1 S Kuu^-1 Kuv)
        Every term is a Kronecker product of per-dimension pieces that need only the main / first off diagonals of
        P_d and Q_d.  Deviations from the reference, on purpose: the last term uses Kuu^-1 S Kuu^-1 (the reference
        writes S^-1 at :942, its 1-D twin Sigma^-1 at gridded_univariate_structure.py:699 -- SURVEY.md appendix B); the
        reference's `[delta, delta]` row (:836) is kept as is.  Returns cells in row-major (n_b0 x n_b0) order."""
        from ... import _lib as L
        from ...params import GriddedMarginals
        plan = self._ensure_plan()
        theta = self._theta()
        m = self.variational_mean.detach().to(torch.float64).contiguous()
        Lc = torch.cat([c.detach().to(torch.float64).reshape(-1) for c in self._chols()]).contiguous()
        plan.grid_forward(theta, m, Lc)
        nb, pad = self.n_b0_splines, self.padding_factor
        alpha = plan.workspace(L.WS_ALPHA).reshape(plan.m_per_dim)
        deltas = [self.b1_basis_1.delta.to(torch.float64).item(), self.b1_basis_2.delta.to(torch.float64).item()]
        sl = lambda a: slice(pad + a, pad + a + nb)
        mean = deltas[0] * deltas[1] * (alpha[sl(0), sl(0)] + alpha[sl(0), sl(1)] + alpha[sl(1), sl(0)] + alpha[sl(1), sl(1)])
        kvv, kpk, kqk = [], [], []
        for d in range(2):
            n = plan.m_per_dim[d]
            P = plan.workspace(L.WS_P, d)
            qb = plan.workspace(L.WS_QBAND, d)
            pd_, po = torch.diagonal(P), torch.diagonal(P, 1)
            qd, qo = qb[:n], qb[n:2 * n - 1]
            i = torch.arange(pad, pad + nb, device=P.device)
            d2 = deltas[d] ** 2
            kpk.append(d2 * (pd_[i] + 2 * po[i] + pd_[i + 1]))
            kqk.append(d2 * (qd[i] + 2 * qo[i] + qd[i + 1]))
            l, s2 = theta[d], theta[2 + d]
            b0d = [self.b0_delta_1, self.b0_delta_2][d].to(torch.float64).to(P.device)
            kvv.append((l ** 2 * s2 * 2 * (torch.exp(-b0d / l) + b0d / l - 1)).expand(nb))
        outer = lambda a, b: (a[:, None] * b[None, :])
        var = outer(kvv[0], kvv[1]) - outer(kpk[0], kpk[1]) + outer(kqk[0], kqk[1])
        return GriddedMarginals(mean.reshape(-1), var.reshape(-1))


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

    def q_v(self, optimal: bool = False):
        """For this model the inducing variables are the gridded cell integrals: q(v) is q(u) (:1409-1433) -- the learned
        one, or with `optimal=True` the reference's closed form."""
        return self.q_u(optimal)
