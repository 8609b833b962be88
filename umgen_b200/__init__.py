"""umgen_b200 -- B200-native next-scene decode engine for UMGen (sm_100a CUDA behind a C-ABI).

Importing the package is light (no CUDA needed); creating an engine loads
``umgen_b200/lib/libumgen_sm100.so`` and raises if it, or a GPU, is missing -- there is no CPU
fallback."""
from .config import ModelConfig, SampleConfig  # noqa: F401

__version__ = "0.1.0"
