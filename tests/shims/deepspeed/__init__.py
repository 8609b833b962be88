"""deepspeed is imported by the reference's module.py for activation checkpointing during training only (TEST INFRASTRUCTURE)."""
from . import checkpointing  # noqa: F401
