"""Parameter containers with the attribute paths the reference's drivers touch
(`model.kernel_1.outputscale`, `.base_kernel.lengthscale`, `model.likelihood.noise`, get and set;
kronecker_structure.py:27-32, 55-61).  gpytorch is not a dependency; the semantics assumed are gpytorch's
defaults: softplus-positive raw parameters initialised at 0, noise = softplus(raw) + 1e-4, shapes (1,1) / () / (1,)."""
import torch
from torch import nn
from torch.nn.functional import softplus

NOISE_LOWER_BOUND = 1e-4


def inv_softplus(x: torch.Tensor) -> torch.Tensor:
    return x + torch.log(-torch.expm1(-x))


def _as_like(value, ref: torch.Tensor) -> torch.Tensor:
    """Python numbers are taken at the parameter's precision (not through a float32 temporary)."""
    if torch.is_tensor(value):
        return value.detach().to(device=ref.device, dtype=ref.dtype)
    return torch.tensor(value, dtype=ref.dtype, device=ref.device)


class MaternKernel(nn.Module):
    def __init__(self, nu: float = 0.5, active_dims=None):
        super().__init__()
        if nu != 0.5:
            raise NotImplementedError("the gridded sparse models use Matern-1/2 only (kronecker_structure.py:14)")
        self.nu = nu
        self.active_dims = active_dims
        self.raw_lengthscale = nn.Parameter(torch.zeros(1, 1))

    @property
    def lengthscale(self) -> torch.Tensor:
        return softplus(self.raw_lengthscale)

    @lengthscale.setter
    def lengthscale(self, value):
        value = _as_like(value, self.raw_lengthscale)
        with torch.no_grad():
            self.raw_lengthscale.copy_(inv_softplus(value).expand_as(self.raw_lengthscale))


class ScaleKernel(nn.Module):
    def __init__(self, base_kernel: nn.Module):
        super().__init__()
        self.base_kernel = base_kernel
        self.raw_outputscale = nn.Parameter(torch.zeros(()))

    @property
    def outputscale(self) -> torch.Tensor:
        return softplus(self.raw_outputscale)

    @outputscale.setter
    def outputscale(self, value):
        value = _as_like(value, self.raw_outputscale)
        with torch.no_grad():
            self.raw_outputscale.copy_(inv_softplus(value).reshape(()))


class HomoskedasticNoise(nn.Module):
    def __init__(self):
        super().__init__()
        self.raw_noise = nn.Parameter(torch.zeros(1))

    @property
    def noise(self) -> torch.Tensor:
        return softplus(self.raw_noise) + NOISE_LOWER_BOUND

    @noise.setter
    def noise(self, value):
        value = _as_like(value, self.raw_noise)
        with torch.no_grad():
            self.raw_noise.copy_(inv_softplus(value - NOISE_LOWER_BOUND).expand_as(self.raw_noise))


class GaussianLikelihood(nn.Module):
    def __init__(self):
        super().__init__()
        self.noise_covar = HomoskedasticNoise()

    @property
    def noise(self) -> torch.Tensor:
        return self.noise_covar.noise

    @noise.setter
    def noise(self, value):
        self.noise_covar.noise = value


class GriddedNormal:
    """MultivariateNormal-like result of q_u(): .mean (M,), Kronecker factors of the covariance, and a dense
    .covariance_matrix built on demand (only sensible for small M)."""

    def __init__(self, mean: torch.Tensor, cov_factors):
        self.mean = mean
        self.loc = mean
        self.cov_factors = list(cov_factors)       # S_d = L_d L_d^T, covariance = kron_d S_d

    @property
    def covariance_matrix(self) -> torch.Tensor:
        out = self.cov_factors[0]
        for S in self.cov_factors[1:]:
            out = torch.kron(out, S)
        return out

    @property
    def variance(self) -> torch.Tensor:
        out = torch.diagonal(self.cov_factors[0])
        for S in self.cov_factors[1:]:
            out = torch.kron(out, torch.diagonal(S))
        return out

    @property
    def stddev(self) -> torch.Tensor:
        return self.variance.sqrt()

    def confidence_region(self):
        s2 = 2 * self.stddev
        return self.mean - s2, self.mean + s2


class DenseNormal:
    """MultivariateNormal-like result of the reference's closed-form (dense, small-M) formulas: .mean, .covariance_matrix."""

    def __init__(self, mean: torch.Tensor, covariance_matrix: torch.Tensor):
        self.mean = mean
        self.loc = mean
        self.covariance_matrix = covariance_matrix

    @property
    def variance(self) -> torch.Tensor:
        return torch.diagonal(self.covariance_matrix)

    @property
    def stddev(self) -> torch.Tensor:
        return self.variance.clamp_min(0).sqrt()

    def confidence_region(self):
        s2 = 2 * self.stddev
        return self.mean - s2, self.mean + s2


class GriddedMarginals:
    """Result of posterior(x*) / posterior_predictive(x*): marginal mean and variance at the test points.  The
    reference returns a dense N* x N* MultivariateNormal (kronecker_structure.py:199-247); only its marginals --
    what the notebooks plot and score -- are computed here."""

    def __init__(self, mean: torch.Tensor, variance: torch.Tensor):
        self.mean = mean
        self.loc = mean
        self.variance = variance

    @property
    def stddev(self) -> torch.Tensor:
        return self.variance.clamp_min(0).sqrt()

    def confidence_region(self):
        s2 = 2 * self.stddev
        return self.mean - s2, self.mean + s2

    @property
    def covariance_matrix(self):
        raise NotImplementedError("only the marginal variances are computed; use .variance")
