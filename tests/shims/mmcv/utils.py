"""mmcv.utils.Registry / build_from_cfg with the semantics the reference relies on (class object or name as cfg['type'])."""


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
    args = dict(cfg)
    t = args.pop("type")
    cls = registry.get(t) if isinstance(t, str) else t
    if cls is None:
        raise KeyError(f"{t} is not in the {registry.name} registry")
    for k, v in (default_args or {}).items():
        args.setdefault(k, v)
    return cls(**args)
