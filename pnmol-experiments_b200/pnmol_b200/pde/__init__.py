"""PDE problem containers and recipes (API of src/pnmol/pde)."""
from . import examples, problems  # noqa: F401
from .problems import Reaction  # noqa: F401
