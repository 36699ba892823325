"""Step-size rules (API of src/pnmol/odetools)."""
from . import step  # noqa: F401
