from .bspline import SplineBasis, B0SplineBasis, B1SplineBasis  # noqa: F401
