"""matplotlib is imported by the reference's visualiser for colour maps it does not use on the evaluation path (TEST INFRASTRUCTURE)."""
from . import pyplot  # noqa: F401
