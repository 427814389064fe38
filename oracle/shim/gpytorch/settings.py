import contextlib


@contextlib.contextmanager
def fast_pred_var(*a, **k):
    yield


@contextlib.contextmanager
def max_cholesky_size(*a, **k):
    yield
