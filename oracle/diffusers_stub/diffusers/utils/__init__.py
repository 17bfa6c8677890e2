"""Stub of diffusers.utils: BaseOutput, logging, WEIGHTS_NAME."""
import logging as _pylog
from collections import OrderedDict
from dataclasses import fields

WEIGHTS_NAME = "diffusion_pytorch_model.bin"


class BaseOutput(OrderedDict):
    def __post_init__(self):
        for f in fields(self):
            v = getattr(self, f.name)
            if v is not None:
                self[f.name] = v

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())

    def __getitem__(self, k):
        if isinstance(k, str):
            return dict(self.items())[k]
        return self.to_tuple()[k]


class logging:  # noqa: N801  (mirrors `diffusers.utils.logging` module usage)
    @staticmethod
    def get_logger(name):
        return _pylog.getLogger(name)


def deprecate(*a, **k):
    pass
