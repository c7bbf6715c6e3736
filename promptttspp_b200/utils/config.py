"""Minimal `_target_` instantiation of the reference's yaml configs -- what `hydra.utils.instantiate` +
OmegaConf interpolation do for app.py:28-39 and egs/proposed/bin/synthesize.py:105-116, without either package.

    cfg = load_yaml("conf/model/prompttts_mdn_v2_wo_erg_final_demo.yaml")
    model = instantiate(cfg)                    # `promptttspp.` targets resolve to `promptttspp_b200.`

Supported (everything the shipped inference configs use): nested dicts/lists, `_target_` classes or callables
instantiated depth-first with their sibling keys as kwargs (Hydra's default `_recursive_=True`), `_partial_`,
relative interpolations `${.a.b}` / `${..a}` / `${...a.b}` (one dot = the container holding the key, each further dot
one level up) and absolute ones `${a.b}`, also inside strings.  A `defaults:` list is resolved by `load_config`
against the conf directory (group: option -> <group>/<option>.yaml mounted under the group key).
"""
import importlib
import re
from pathlib import Path
from typing import Any, Dict, Mapping, Optional

import yaml

DEFAULT_PREFIX_MAP = {"promptttspp.": "promptttspp_b200."}
_INTERP = re.compile(r"\$\{([^${}]+)\}")


def load_yaml(path) -> Any:
    with open(path, "r") as f:
        return yaml.safe_load(f)


def load_config(path, groups: Optional[Mapping[str, str]] = None) -> Dict[str, Any]:
    """A top-level config with a Hydra `defaults:` list (demo.yaml / synthesize.yaml): every `group: option` entry
    mounts conf/<group>/<option>.yaml under `group`; `groups` overrides options; `_self_`, `override hydra/...` and the
    `hydra:` block are ignored."""
    path = Path(path)
    cfg = load_yaml(path) or {}
    defaults = cfg.pop("defaults", [])
    cfg.pop("hydra", None)
    out: Dict[str, Any] = {}
    for d in defaults:
        if not isinstance(d, dict):
            continue
        for group, option in d.items():
            if group.startswith("override ") or option is None:
                continue
            option = (groups or {}).get(group, option)
            sub = path.parent / group / f"{option}.yaml"
            if sub.exists():
                out[group] = load_yaml(sub)
    out.update(cfg)
    return out


def _lookup(root, path_keys):
    node = root
    for k in path_keys:
        if isinstance(node, list):
            node = node[int(k)]
        else:
            node = node[k]
    return node


def _resolve_value(root, container_path, text, depth=0):
    """Resolve the interpolations inside the string `text`, which sits in the container at `container_path`."""
    if depth > 32:
        raise ValueError(f"interpolation cycle at {'.'.join(map(str, container_path))}: {text}")

    def target(expr):
        expr = expr.strip()
        if expr.startswith("."):
            dots = len(expr) - len(expr.lstrip("."))
            base = list(container_path[: len(container_path) - (dots - 1)]) if dots > 1 else list(container_path)
            if dots - 1 > len(container_path):
                raise KeyError(f"interpolation {expr!r} climbs above the root")
            keys = base + [k for k in expr.lstrip(".").split(".") if k]
        else:
            keys = expr.split(".")
        val = _lookup(root, keys)
        if isinstance(val, str) and _INTERP.search(val):
            val = _resolve_value(root, keys[:-1], val, depth + 1)
        return val

    m = _INTERP.fullmatch(text.strip())
    if m:  # the whole value is one interpolation: keep the referenced type
        return target(m.group(1))
    return _INTERP.sub(lambda mm: str(target(mm.group(1))), text)


def resolve(cfg):
    """Return a copy of `cfg` with every `${...}` interpolation replaced by the value it refers to."""

    def walk(node, path):
        if isinstance(node, dict):
            return {k: walk(v, path + [k]) for k, v in node.items()}
        if isinstance(node, list):
            return [walk(v, path + [i]) for i, v in enumerate(node)]
        if isinstance(node, str) and _INTERP.search(node):
            return _resolve_value(cfg, path[:-1], node)
        return node

    return walk(cfg, [])


def locate(target: str, prefix_map: Optional[Mapping[str, str]] = None):
    """'pkg.mod.Class' -> the object, after re-pointing prefixes (reference package -> this package)."""
    for old, new in (DEFAULT_PREFIX_MAP if prefix_map is None else prefix_map).items():
        if target.startswith(old):
            target = new + target[len(old):]
            break
    parts = target.split(".")
    for cut in range(len(parts) - 1, 0, -1):
        try:
            obj = importlib.import_module(".".join(parts[:cut]))
        except ModuleNotFoundError:
            continue
        for name in parts[cut:]:
            obj = getattr(obj, name)
        return obj
    raise ImportError(f"cannot locate {target!r}")


def instantiate(cfg, prefix_map: Optional[Mapping[str, str]] = None, _resolved=False, **overrides):
    """Build the object a `_target_` node describes; plain containers are returned with their children built."""
    if not _resolved:
        cfg = resolve(cfg)
    if isinstance(cfg, list):
        return [instantiate(v, prefix_map, True) for v in cfg]
    if not isinstance(cfg, dict):
        return cfg
    kwargs = {k: instantiate(v, prefix_map, True) for k, v in cfg.items() if k not in ("_target_", "_partial_")}
    if "_target_" not in cfg:
        return kwargs
    kwargs.update(overrides)
    fn = locate(cfg["_target_"], prefix_map)
    if cfg.get("_partial_"):
        import functools

        return functools.partial(fn, **kwargs)
    return fn(**kwargs)
