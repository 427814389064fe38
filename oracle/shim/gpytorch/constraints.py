"""Positive / GreaterThan constraints (softplus transform), as in gpytorch.constraints."""
import torch
from torch.nn.functional import softplus


def inv_softplus(x):
    return x + torch.log(-torch.expm1(-x))


class GreaterThan:
    def __init__(self, lower_bound):
        self.lower_bound = float(lower_bound)

    def transform(self, raw):
        return softplus(raw) + self.lower_bound

    def inverse_transform(self, value):
        return inv_softplus(value - self.lower_bound)


class Positive(GreaterThan):
    def __init__(self):
        super().__init__(0.0)
