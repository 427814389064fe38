from . import Module


class GP(Module):
    pass


class ExactGP(GP):
    """Off the hot path; only here so `src/models/exact/*` can be imported by gridded_* modules."""

    def __init__(self, train_inputs=None, train_targets=None, likelihood=None):
        super().__init__()
        self.train_inputs = (train_inputs,) if train_inputs is not None and not isinstance(train_inputs, tuple) else train_inputs
        self.train_targets = train_targets
        self.likelihood = likelihood
