"""ScaleKernel / MaternKernel / ProductKernel subset (kronecker_structure.py:30-32, 267)."""
import math
import torch
from torch import nn
from . import Module, _DenseLazy
from .constraints import Positive


class Kernel(Module):
    def __init__(self, active_dims=None):
        super().__init__()
        self.active_dims = None if active_dims is None else torch.as_tensor(active_dims, dtype=torch.long)

    def _select(self, x):
        if x.dim() == 1:
            x = x.unsqueeze(-1)
        if self.active_dims is not None:
            x = x.index_select(-1, self.active_dims)
        return x

    def __call__(self, x1, x2=None):
        x1_ = self._select(x1)
        x2_ = x1_ if x2 is None else self._select(x2)
        return _DenseLazy(self.forward(x1_, x2_))

    def __mul__(self, other):
        return ProductKernel(self, other)


class MaternKernel(Kernel):
    def __init__(self, nu=2.5, active_dims=None):
        super().__init__(active_dims)
        self.nu = nu
        self.raw_lengthscale = nn.Parameter(torch.zeros(1, 1))
        self.raw_lengthscale_constraint = Positive()

    @property
    def lengthscale(self):
        return self.raw_lengthscale_constraint.transform(self.raw_lengthscale)

    @lengthscale.setter
    def lengthscale(self, value):
        value = torch.as_tensor(value).to(self.raw_lengthscale)
        self.raw_lengthscale.data.copy_(
            self.raw_lengthscale_constraint.inverse_transform(value).expand_as(self.raw_lengthscale))

    def forward(self, x1, x2):
        # exact pairwise distances (gpytorch's matmul-based sq_dist leaves ~1e-8 noise on the diagonal of
        # k(X, X) while training; that noise is not part of the reference's algorithm and is not reproduced)
        diff = (x1.unsqueeze(-2) - x2.unsqueeze(-3)) / self.lengthscale
        d = diff.pow(2).sum(-1).clamp_min(1e-30).sqrt()
        if self.nu == 0.5:
            return torch.exp(-d)
        if self.nu == 1.5:
            return (1 + math.sqrt(3) * d) * torch.exp(-math.sqrt(3) * d)
        if self.nu == 2.5:
            return (1 + math.sqrt(5) * d + 5.0 / 3.0 * d ** 2) * torch.exp(-math.sqrt(5) * d)
        raise NotImplementedError


class ScaleKernel(Kernel):
    def __init__(self, base_kernel):
        super().__init__(None)
        self.base_kernel = base_kernel
        self.raw_outputscale = nn.Parameter(torch.zeros(()))
        self.raw_outputscale_constraint = Positive()

    @property
    def outputscale(self):
        return self.raw_outputscale_constraint.transform(self.raw_outputscale)

    @outputscale.setter
    def outputscale(self, value):
        value = torch.as_tensor(value).to(self.raw_outputscale)
        self.raw_outputscale.data.copy_(self.raw_outputscale_constraint.inverse_transform(value))

    def __call__(self, x1, x2=None):
        return _DenseLazy(self.base_kernel(x1, x2).tensor * self.outputscale)


class ProductKernel(Kernel):
    def __init__(self, *kernels):
        super().__init__(None)
        self.kernels = nn.ModuleList(kernels)

    def __call__(self, x1, x2=None):
        out = None
        for k in self.kernels:
            v = k(x1, x2).tensor
            out = v if out is None else out * v
        return _DenseLazy(out)
