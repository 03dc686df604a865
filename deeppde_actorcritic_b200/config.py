"""Attribute-style access over the JSON config (the reference uses ``munch.munchify``, main.py:33)."""
from __future__ import annotations

import json


class Config(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return v

    def __setattr__(self, k, v):
        self[k] = v


def munchify(obj):
    if isinstance(obj, dict):
        return Config((k, munchify(v)) for k, v in obj.items())
    if isinstance(obj, list):
        return [munchify(v) for v in obj]
    return obj


def load_config(path):
    with open(path) as f:
        return munchify(json.load(f))
