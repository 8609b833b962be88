"""Minimal stand-in for the parts of mmcv the reference's evaluate.py / config / plugin files use (TEST INFRASTRUCTURE)."""
import os
import runpy
import types

from . import utils  # noqa: F401
from .utils import Registry, build_from_cfg  # noqa: F401


class Config(dict):
    """mmcv.Config.fromfile for a python config: the file's public, non-module, non-callable-module globals as a dict with attribute access."""

    @staticmethod
    def fromfile(path):
        ns = runpy.run_path(os.path.abspath(path))
        cfg = Config()
        for k, v in ns.items():
            if k.startswith("__") or isinstance(v, types.ModuleType) or isinstance(v, (types.FunctionType, type)):
                continue
            cfg[k] = v
        return cfg

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def imfrombytes(*a, **k):
    raise NotImplementedError("image loading is outside the evaluation path")
