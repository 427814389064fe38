"""Minimal stand-in for the `gpytorch` symbols the reference model files use.

TEST INFRASTRUCTURE ONLY (see oracle/shim/README.md).  Used by oracle/make_golden.py so that
/root/reference/src/models/sparse/*.py import unmodified.  Call sites being served:
kronecker_structure.py:15,27-32,101-103,170-172,265-273 ; gridded_kronecker_structure.py:906-916 ;
univariate_structure.py:15,41-42,243-258.
"""
import math
import torch
from torch import nn
from torch.nn.functional import softplus


def _inv_softplus(x):
    return x + torch.log(-torch.expm1(-x))


class Module(nn.Module):
    """gpytorch.Module: an nn.Module with `.initialize(**kwargs)`."""

    def initialize(self, **kwargs):
        for name, val in kwargs.items():
            if not torch.is_tensor(val):
                val = torch.as_tensor(val)
            if hasattr(type(self), name) and isinstance(getattr(type(self), name), property):
                setattr(self, name, val)
            else:
                p = getattr(self, name)
                p.data.copy_(val.to(p).expand_as(p))
        return self


class _DenseLazy:
    """Eager dense matrix posing as a LinearOperator / LazyTensor."""

    def __init__(self, tensor):
        self.tensor = tensor

    # --- LinearOperator API subset ---
    def inv_matmul(self, rhs):
        L = torch.linalg.cholesky(self.tensor)
        if rhs.dim() == 1:
            return torch.cholesky_solve(rhs.unsqueeze(-1), L).squeeze(-1)
        return torch.cholesky_solve(rhs, L)

    def evaluate(self):
        return self.tensor

    def to_dense(self):
        return self.tensor

    def add_diagonal(self, diag):
        return _DenseLazy(self.tensor + torch.diag_embed(diag.to(self.tensor.dtype)))

    def add_low_rank(self, low_rank_mat):
        # LinearOperator.add_low_rank: A + B B^T (call site: kronecker_structure.py:462, DiagLinearOperator(alpha).add_low_rank(beta))
        return _DenseLazy(self.tensor + low_rank_mat @ low_rank_mat.transpose(-1, -2))

    def mul(self, other):
        if isinstance(other, _DenseLazy):
            other = other.tensor
        if torch.is_tensor(other) and other.numel() == 1:
            other = other.reshape(())
        return _DenseLazy(self.tensor.mul(other))

    def __add__(self, other):
        if isinstance(other, _DenseLazy):
            other = other.tensor
        return _DenseLazy(self.tensor + other)

    __radd__ = __add__

    def __matmul__(self, other):
        if isinstance(other, _DenseLazy):
            other = other.tensor
        return self.tensor @ other

    def __rmatmul__(self, other):
        return other @ self.tensor

    @property
    def shape(self):
        return self.tensor.shape

    @property
    def dtype(self):
        return self.tensor.dtype

    @property
    def T(self):
        return _DenseLazy(self.tensor.T)

    def size(self, *a):
        return self.tensor.size(*a)


def lazify(obj):
    if isinstance(obj, _DenseLazy):
        return obj
    return _DenseLazy(obj)


from . import kernels, likelihoods, means, distributions, settings, models, mlls, constraints  # noqa: E402,F401
