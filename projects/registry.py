"""MODELS / DATASETS registries (reference projects/registry.py:1-3).

The reference builds them with ``mmcv.utils.Registry``; when mmcv is importable it is used as-is, otherwise a
minimal equivalent with the two entry points the reference uses -- the ``register_module()`` decorator and
``build_from_cfg(cfg, registry)`` accepting a class object or a registered name as ``cfg["type"]``."""
try:                                            # pragma: no cover - mmcv is not installed in the build image
    from mmcv.utils import Registry, build_from_cfg
except Exception:
    class Registry:
        def __init__(self, name):
            self._name, self._module_dict = name, {}

        @property
        def name(self):
            return self._name

        @property
        def module_dict(self):
            return self._module_dict

        def get(self, key):
            return self._module_dict.get(key)

        def register_module(self, name=None, force=False, module=None):
            def _register(cls):
                key = name or cls.__name__
                if key in self._module_dict and not force:
                    raise KeyError(f"{key} is already registered in {self._name}")
                self._module_dict[key] = cls
                return cls
            return _register(module) if module is not None else _register

    def build_from_cfg(cfg, registry, default_args=None):
        if not isinstance(cfg, dict) or "type" not in cfg:
            raise TypeError("cfg must be a dict containing the key 'type'")
        args = dict(cfg)
        obj_type = args.pop("type")
        if isinstance(obj_type, str):
            cls = registry.get(obj_type)
            if cls is None:
                raise KeyError(f"{obj_type} is not in the {registry.name} registry")
        else:
            cls = obj_type
        for k, v in (default_args or {}).items():
            args.setdefault(k, v)
        return cls(**args)

MODELS = Registry("models")
DATASETS = Registry("datasets")
