"""ToeplitzLinearOperator / DiagLinearOperator subset, eager and dense.
Call sites: gridded_kronecker_structure.py:738,747,754,883,1321 ; kronecker_structure.py:567,576,583,737."""
import torch
from gpytorch import _DenseLazy


class ToeplitzLinearOperator(_DenseLazy):
    def __init__(self, column):
        n = column.shape[-1]
        idx = (torch.arange(n)[:, None] - torch.arange(n)[None, :]).abs()
        super().__init__(column[idx])


class DiagLinearOperator(_DenseLazy):
    def __init__(self, diag):
        super().__init__(torch.diag_embed(diag))
