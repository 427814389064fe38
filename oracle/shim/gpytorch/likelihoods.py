"""GaussianLikelihood subset (kronecker_structure.py:27,61,146; gridded_kronecker_structure.py:906)."""
import torch
from torch import nn
from . import Module
from .constraints import GreaterThan


class HomoskedasticNoise(Module):
    def __init__(self):
        super().__init__()
        self.raw_noise = nn.Parameter(torch.zeros(1))
        self.raw_noise_constraint = GreaterThan(1e-4)

    @property
    def noise(self):
        return self.raw_noise_constraint.transform(self.raw_noise)

    @noise.setter
    def noise(self, value):
        value = torch.as_tensor(value).to(self.raw_noise)
        self.raw_noise.data.copy_(self.raw_noise_constraint.inverse_transform(value).expand_as(self.raw_noise))


class GaussianLikelihood(Module):
    def __init__(self):
        super().__init__()
        self.noise_covar = HomoskedasticNoise()

    @property
    def noise(self):
        return self.noise_covar.noise

    @noise.setter
    def noise(self, value):
        self.noise_covar.noise = value

    def __call__(self, dist):
        from .distributions import MultivariateNormal
        cov = dist.covariance_matrix
        eye = torch.eye(cov.shape[-1], dtype=cov.dtype)
        return MultivariateNormal(dist.mean, cov + self.noise * eye)
