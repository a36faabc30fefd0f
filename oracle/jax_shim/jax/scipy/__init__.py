"""jax.scipy stand-in (SciPy). TEST INFRASTRUCTURE ONLY."""
from . import linalg, special  # noqa: F401
