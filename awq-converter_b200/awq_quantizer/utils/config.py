"""YAML configuration with the reference's schema and accessors (utils/config.py:11-177): defaults from
config/default_config.yaml, deep-merged user file, dotted get/set, hub-id heuristics for model.path."""
from __future__ import annotations

import os
from typing import Any, Dict, Optional

import yaml

_DEFAULT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "config", "default_config.yaml")


def _merge(dst: Dict[str, Any], src: Dict[str, Any]) -> Dict[str, Any]:
    for k, v in (src or {}).items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v
    return dst


class Config:
    def __init__(self, config_path: Optional[str] = None):
        self.config: Dict[str, Any] = {}
        if os.path.exists(_DEFAULT):
            with open(_DEFAULT, "r", encoding="utf-8") as f:
                self.config = yaml.safe_load(f) or {}
        if config_path is not None and os.path.exists(config_path):
            with open(config_path, "r", encoding="utf-8") as f:
                _merge(self.config, yaml.safe_load(f) or {})
        m = self.config.setdefault("model", {})
        for k, d in (("path", ""), ("hub_model_id", ""), ("from_hub", False), ("revision", "main"), ("token", None)):
            m.setdefault(k, d)
        if m["hub_model_id"]:
            m["path"], m["from_hub"] = m["hub_model_id"], True
        elif m["path"] and "/" in m["path"] and not os.path.exists(m["path"]):
            m["hub_model_id"], m["from_hub"] = m["path"], True

    def get(self, key: str, default: Any = None) -> Any:
        node: Any = self.config
        for part in key.split("."):
            if isinstance(node, dict) and part in node:
                node = node[part]
            else:
                return default
        return node

    def set(self, key: str, value: Any) -> None:
        parts = key.split(".")
        node = self.config
        for part in parts[:-1]:
            node = node.setdefault(part, {})
        node[parts[-1]] = value

    def save(self, path: str) -> None:
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        with open(path, "w", encoding="utf-8") as f:
            yaml.dump(self.config, f, default_flow_style=False)

    __getitem__ = get

    def __setitem__(self, key: str, value: Any) -> None:
        self.set(key, value)


def load_config(config_path: Optional[str] = None) -> Config:
    return Config(config_path)
