"""State-space building blocks (API of src/pnmol/base)."""
from . import iwp, rv, sqrt, stacked_ssm  # noqa: F401
