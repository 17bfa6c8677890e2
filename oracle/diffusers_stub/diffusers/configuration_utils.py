"""Stub of diffusers.configuration_utils: ConfigMixin / register_to_config / FrozenDict."""
import functools
import inspect


class FrozenDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class ConfigMixin:
    config_name = "config.json"

    def register_to_config(self, **kw):
        cfg = dict(getattr(self, "_internal_dict", {}))
        cfg.update(kw)
        self._internal_dict = FrozenDict(cfg)

    @property
    def config(self):
        return self._internal_dict

    @classmethod
    def from_config(cls, config, **kwargs):
        sig = inspect.signature(cls.__init__).parameters
        init = {k: v for k, v in dict(config).items() if k in sig and k != "self"}
        init.update({k: v for k, v in kwargs.items() if k in sig})
        return cls(**init)


def register_to_config(init):
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        sig = inspect.signature(init)
        params = list(sig.parameters.items())[1:]
        cfg = {n: p.default for n, p in params if p.default is not inspect.Parameter.empty}
        for (n, _), a in zip(params, args):
            cfg[n] = a
        cfg.update(kwargs)
        init(self, *args, **kwargs)
        self.register_to_config(**cfg)
    return inner
