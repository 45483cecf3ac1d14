"""Minimal stand-in for omegaconf.DictConfig + Hydra overrides (neither is installed here).

The reference hands ``cfg.algorithm`` (an OmegaConf DictConfig) to ``DFoTVideo.__init__``; every key the
sampling path reads is listed in SURVEY.md §5.  ``to_config`` accepts a plain dict, one of these objects,
or a real omegaconf DictConfig (converted with OmegaConf.to_container when omegaconf is importable).
"""
import copy
from typing import Any


class DictConfig(dict):
    """attr-style access to a nested dict (``cfg.diffusion.timesteps``), `.get` as in OmegaConf."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key) from None

    def __setattr__(self, key, value):
        self[key] = value

    def __deepcopy__(self, memo):
        return DictConfig({k: copy.deepcopy(v, memo) for k, v in self.items()})


def to_config(obj: Any) -> Any:
    if isinstance(obj, DictConfig):
        return obj
    if isinstance(obj, dict):
        return DictConfig({k: to_config(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return [to_config(v) for v in obj]
    if type(obj).__module__.startswith("omegaconf"):
        from omegaconf import OmegaConf  # type: ignore
        return to_config(OmegaConf.to_container(obj, resolve=True))
    return obj


def to_container(obj: Any) -> Any:
    """OmegaConf.to_container(resolve=True) equivalent: plain dict / list tree."""
    if isinstance(obj, dict):
        return {k: to_container(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [to_container(v) for v in obj]
    if type(obj).__module__.startswith("omegaconf"):
        from omegaconf import OmegaConf  # type: ignore
        return OmegaConf.to_container(obj, resolve=True)
    return obj


def apply_overrides(cfg: dict, overrides) -> dict:
    """Hydra-style ``a.b.c=value`` / ``+a.b=value`` / ``++a.b=value`` overrides on a plain tree (YAML-typed values)."""
    import yaml
    cfg = copy.deepcopy(to_container(cfg))
    for item in overrides:
        key, _, raw = item.partition("=")
        key = key.lstrip("+")
        node = cfg
        parts = key.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = yaml.safe_load(raw)
    return cfg
