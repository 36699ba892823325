"""State-space building blocks (API of src/pnmol/base)."""
from . import iwp, kalman, rv, sqrt, stacked_ssm  # noqa: F401
