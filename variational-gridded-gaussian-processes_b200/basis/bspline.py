"""B-spline bases with the reference's container interface (src/basis/bspline.py:81-112): attributes
`.mesh .m .delta .n_basis_functions`, `__call__(x) -> (n_basis_functions, N)`.

Evaluation runs in libvggp (`vggp_features_dense` -> the same device stencil function the fused ELBO kernel uses),
never through a Python loop over basis objects (bspline.py:92-94) and never on the CPU."""
import torch

from .. import _lib
from ..plan import GridPlan


class SplineBasis:
    order = None

    def __init__(self, mesh: torch.Tensor):
        self.mesh = mesh
        self.m = mesh.size(0) - (self.order + 1)
        self.delta = mesh[1] - mesh[0]
        self._plans = {}

    def _plan(self, dtype, device) -> GridPlan:
        key = (dtype, str(device))
        if key not in self._plans:
            self._plans[key] = GridPlan(_lib.B1_ASVGP, [self.mesh], dtype, device)
        return self._plans[key]


class B0SplineBasis(SplineBasis):
    """Order-0 (indicator) basis: only `.mesh/.m/.delta` are used on the ELBO path; calling
    it evaluates the closed-interval indicators of bspline.py:15-20 on the device."""
    order = 0

    def __init__(self, mesh: torch.Tensor):
        super().__init__(mesh)
        self.n_basis_functions = self.m

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("basis evaluation needs CUDA tensors: this package has no CPU path")
        mesh = self.mesh.to(x.device).to(x.dtype)
        return (torch.logical_and(x[None, :] >= mesh[:-1, None], x[None, :] <= mesh[1:, None]) * 1)


class B1SplineBasis(SplineBasis):
    """Order-1 (hat) basis: left half hat + K-2 hats + right half hat = K functions (bspline.py:106-112)."""
    order = 1

    def __init__(self, mesh: torch.Tensor):
        super().__init__(mesh)
        self.n_basis_functions = mesh.size(0)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("basis evaluation needs CUDA tensors: this package has no CPU path")
        return self._plan(x.dtype, x.device).features_dense(0, x.reshape(-1))

    def stencil(self, x: torch.Tensor):
        """(c, w_lo, w_hi): the two non-zeros of each column of __call__(x); c = -1 outside the mesh."""
        return self._plan(x.dtype, x.device).b1_stencil(0, x.reshape(-1))
