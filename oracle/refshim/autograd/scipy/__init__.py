"""TEST INFRASTRUCTURE ONLY.  `autograd.scipy` stand-in."""
from . import linalg, special, stats  # noqa: F401
