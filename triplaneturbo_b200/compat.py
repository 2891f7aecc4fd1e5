"""Minimal stand-ins for the threestudio plumbing the two plugins rely on: the name registry
(threestudio/__init__.py:1-29), structured config parsing (threestudio/utils/config.py:126-128), the
``BaseModule``/``Updateable`` protocol (threestudio/utils/base.py) and scheduled scalars ``C``
(threestudio/utils/misc.py:66-103).  No omegaconf / lightning dependency: configs are dataclasses filled from
plain dicts (OmegaConf DictConfig objects work too, they are Mapping-like).
"""
import dataclasses
import math
from collections.abc import Mapping
from typing import Any, Optional

import torch.nn as nn

__modules__ = {}


def register(name):
    def decorator(cls):
        __modules__[name] = cls
        return cls
    return decorator


def find(name):
    if name not in __modules__:
        raise KeyError(f"module '{name}' is not registered; known: {sorted(__modules__)}")
    return __modules__[name]


def parse_structured(fields: Any, cfg: Optional[Any] = None) -> Any:
    """dict / Mapping / dataclass instance -> instance of the dataclass ``fields`` (unknown keys are an error,
    like OmegaConf.structured)."""
    if cfg is None:
        return fields()
    if dataclasses.is_dataclass(cfg) and not isinstance(cfg, type):
        cfg = dataclasses.asdict(cfg)
    if not isinstance(cfg, Mapping):
        cfg = dict(cfg)
    names = {f.name for f in dataclasses.fields(fields)}
    unknown = [k for k in cfg.keys() if k not in names]
    if unknown:
        raise KeyError(f"{fields.__qualname__}: unknown config keys {unknown}")
    return fields(**{k: cfg[k] for k in cfg.keys()})


def C(value: Any, epoch: int, global_step: int, interpolation="linear") -> float:
    """Scheduled scalar ``[start_step, start_value, end_value, end_step]`` (threestudio/utils/misc.py:66-103)."""
    if isinstance(value, (int, float)):
        return value
    value = list(value)
    if len(value) == 3:
        value = [0] + value
    if len(value) >= 6:
        select_i = 3
        for i in range(3, len(value) - 2, 2):
            if global_step >= value[i]:
                select_i = i + 2
        if select_i != 3:
            start_value, start_step = value[select_i - 3], value[select_i - 2]
        else:
            start_step, start_value = value[:2]
        end_value, end_step = value[select_i - 1], value[select_i]
        value = [start_step, start_value, end_value, end_step]
    if len(value) != 4:
        raise TypeError(f"scalar specification must have 3, 4 or >= 6 entries, got {value}")
    start_step, start_value, end_value, end_step = value
    current = global_step if isinstance(end_step, int) else epoch
    t = max(min(1.0, (current - start_step) / (end_step - start_step)), 0.0)
    if interpolation == "linear":
        return start_value + (end_value - start_value) * t
    if interpolation == "exp":
        return math.exp(math.log(start_value) * (1 - t) + math.log(end_value) * t)
    raise ValueError(f"unknown interpolation {interpolation}")


class Updateable:
    def do_update_step(self, epoch: int, global_step: int, on_load_weights: bool = False):
        for attr in self.__dir__():
            if attr.startswith("_"):
                continue
            try:
                module = getattr(self, attr)
            except Exception:
                continue
            if isinstance(module, Updateable):
                module.do_update_step(epoch, global_step, on_load_weights=on_load_weights)
        self.update_step(epoch, global_step, on_load_weights=on_load_weights)

    def update_step(self, epoch: int, global_step: int, on_load_weights: bool = False):
        pass


class BaseModule(nn.Module, Updateable):
    @dataclasses.dataclass
    class Config:
        pass

    cfg: Config

    def __init__(self, cfg=None, *args, **kwargs) -> None:
        super().__init__()
        self.cfg = parse_structured(self.Config, cfg)
        self.configure(*args, **kwargs)

    def configure(self, *args, **kwargs) -> None:
        pass

    @property
    def device(self):
        for p in self.parameters():
            return p.device
        for b in self.buffers():
            return b.device
        import torch
        return torch.device("cuda" if torch.cuda.is_available() else "cpu")
