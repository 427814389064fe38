"""MultivariateNormal subset: .mean .covariance_matrix .log_prob (dense Cholesky semantics)."""
import math
import torch
from . import _DenseLazy


class MultivariateNormal:
    def __init__(self, mean, covariance_matrix):
        if isinstance(covariance_matrix, _DenseLazy):
            covariance_matrix = covariance_matrix.tensor
        self.mean = mean
        self.loc = mean
        self.covariance_matrix = covariance_matrix

    @property
    def variance(self):
        return torch.diagonal(self.covariance_matrix, dim1=-2, dim2=-1)

    @property
    def stddev(self):
        return self.variance.sqrt()

    def confidence_region(self):
        s2 = self.stddev * 2
        return self.mean - s2, self.mean + s2

    def log_prob(self, value):
        diff = (value - self.mean).unsqueeze(-1)
        L = torch.linalg.cholesky(self.covariance_matrix)
        sol = torch.cholesky_solve(diff, L)
        quad = (diff * sol).sum()
        logdet = 2.0 * torch.log(torch.diagonal(L)).sum()
        n = diff.shape[0]
        return -0.5 * (quad + logdet + n * math.log(2 * math.pi))
