"""config.yaml loader.  The reference reads its single config file with ``OmegaConf.load`` at every
constructor (model.py:23, transformer.py:13); omegaconf is not a dependency here, PyYAML is
enough because the file is only ever read.  Returned nodes allow both ``cfg.a.b`` and ``cfg["a"]["b"]``
and ``**cfg.node`` unpacking, which is all the reference API surface uses."""
from __future__ import annotations

import os
from typing import Any

import yaml


class ConfigNode(dict):
    def __getattr__(self, k: str) -> Any:
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k: str, v: Any) -> None:
        self[k] = v


def _wrap(o: Any) -> Any:
    if isinstance(o, dict):
        return ConfigNode({k: _wrap(v) for k, v in o.items()})
    if isinstance(o, list):
        return [_wrap(v) for v in o]
    return o


DEFAULT_CONFIG_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config.yaml")


def load_config(path=None) -> ConfigNode:
    with open(path or DEFAULT_CONFIG_PATH) as f:
        return _wrap(yaml.safe_load(f))
