"""Composer for the reference's Hydra configuration tree — the command-line surface of `python -m main ...`
(`configurations/config.yaml:2-7`, `utils/hydra_utils.py:43-112`, `main.py:48-57`) without Hydra / OmegaConf, which
the sampling path does not otherwise need.  Given the user's `configurations/` directory and the reference's own
argument list it returns the resolved tree, whose `algorithm` node is what `DFoTVideo(cfg)` / `DFoTVideoPose(cfg)` take:

    cfg = compose("/path/to/diffusion-forcing-transformer/configurations",
                  ["dataset=realestate10k_mini", "algorithm=dfot_video_pose", "experiment=video_generation",
                   "@diffusion/continuous", "dataset.context_length=1", "dataset.n_frames=8",
                   "algorithm.tasks.prediction.history_guidance.name=vanilla",
                   "+algorithm.tasks.prediction.history_guidance.guidance_scale=4.0"])
    algo = DFoTVideoPose(cfg["algorithm"])

Implemented subset (everything the reference's tree uses):
  * defaults lists, depth first, the file's own body merged last (`_self_` may move it): plain names (a sibling file),
    `group: choice`, `optional group: choice`, `override group: choice`, relative groups with an explicit package
    (`../algorithm/backbone@algorithm.backbone: dit3d`), `null` choices, choices written as interpolations of other
    choices (`${dataset}_${experiment}`);
  * `# @package _global_` headers; default package = the group path;
  * command line: `group=choice`, `key=value`, `+key=value`, `++key=value`, `~key`, and the reference's pre-processing
    of `@shortcut`, `algorithm/backbone=name`, `algorithm/vae=name` (hydra_utils.unwrap_shortcuts);
  * `${a.b.c}` interpolation (whole-value: typed; inside a string: substituted), resolved lazily with cycle detection;
    resolvers Hydra supplies at run time (`${now:...}`, `${hydra:...}`, `${hydra.*}`, `${oc.*}`) are left as they are;
  * OmegaConf's YAML float grammar (`1e-4` is a float), `???` kept as the missing-value marker.
"""
import copy
import os
import re
from typing import Any, Dict, List, Optional, Tuple

import yaml

MISSING = "???"
_FLOAT = re.compile(r"""^(?:[-+]?(?:[0-9][0-9_]*)\.[0-9_]*(?:[eE][-+]?[0-9]+)?
                        |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
                        |\.[0-9_]+(?:[eE][-+][0-9]+)?
                        |[-+]?\.(?:inf|Inf|INF)|\.(?:nan|NaN|NAN))$""", re.X)
_INTERP = re.compile(r"\$\{([^${}]+)\}")


class _Loader(yaml.SafeLoader):
    """PyYAML with OmegaConf's float resolver: YAML 1.1 reads `1e-4` as a string, the reference's tree relies on floats."""


_Loader.add_implicit_resolver("tag:yaml.org,2002:float", _FLOAT, list("-+0123456789."))


def _load_yaml(text: str):
    return yaml.load(text, Loader=_Loader)


def parse_value(text: str) -> Any:
    """A command-line value: YAML flow syntax (`[validation]`, `{a: 1}`, `null`, `true`, numbers), else the string."""
    try:
        return _load_yaml(text)
    except yaml.YAMLError:
        return text


# ------------------------------------------------------------------------------------------ dict helpers
def _merge(dst: dict, src: dict) -> dict:
    """Deep merge `src` into `dst` (dicts merge, everything else — lists included — is replaced)."""
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)
    return dst


def _at_package(body: dict, package: str) -> dict:
    out = body
    for key in reversed([p for p in package.split(".") if p]):
        out = {key: out}
    return out


def _get(tree: dict, path: str):
    node = tree
    for part in path.split("."):
        m = re.fullmatch(r"([^\[\]]+)((?:\[\d+\])*)", part)
        if m is None or not isinstance(node, dict) or m.group(1) not in node:
            raise KeyError(path)
        node = node[m.group(1)]
        for idx in re.findall(r"\[(\d+)\]", m.group(2)):
            node = node[int(idx)]
    return node


def _set(tree: dict, path: str, value, must_exist: Optional[bool]) -> None:
    """must_exist True: plain `a.b=v` (the key has to be there); False: `+a.b=v` (has to be new); None: `++` (either)."""
    parts = path.split(".")
    node = tree
    for p in parts[:-1]:
        if not isinstance(node.get(p), dict):
            if must_exist:
                raise KeyError(f"Could not override '{path}': key '{p}' is not in the config (use +{path}=... to add it)")
            node[p] = {}
        node = node[p]
    leaf = parts[-1]
    if must_exist is True and leaf not in node:
        raise KeyError(f"Could not override '{path}': no such key (use +{path}=... to add it)")
    if must_exist is False and leaf in node:
        raise KeyError(f"Could not append '{path}': the key already exists (use {path}=... or ++{path}=...)")
    if isinstance(value, dict) and isinstance(node.get(leaf), dict):
        _merge(node[leaf], value)
    else:
        node[leaf] = value


# ------------------------------------------------------------------------------------------ shortcuts (hydra_utils.py)
def _yaml_to_cli(path: str, prefix: Optional[str] = None) -> List[str]:
    """hydra_utils._yaml_to_cli: every top-level key of the file becomes a `++key=value` override."""
    with open(path) as f:
        tree = _load_yaml(f.read()) or {}
    return [f"++{prefix + '.' if prefix else ''}{k}=" + yaml.safe_dump(v, default_flow_style=True).strip().removesuffix("...").strip()
            for k, v in tree.items()]


def unwrap_shortcuts(argv: List[str], config_dir: str, config_name: str = "config") -> List[str]:
    """hydra_utils.unwrap_shortcuts:43-98: `@name` -> the overrides of shortcut/name/base.yaml (+ shortcut/name/<dataset>.yaml)
    or of shortcut/name.yaml; `algorithm/backbone=x` / `algorithm/vae=x` -> reset the node and re-fill it from the file."""
    with open(os.path.join(config_dir, f"{config_name}.yaml")) as f:
        defaults = (_load_yaml(f.read()) or {}).get("defaults", [])
    dataset = next((d["dataset"] for d in defaults if isinstance(d, dict) and "dataset" in d), None)
    for arg in argv:
        if arg.startswith("dataset="):
            dataset = arg.split("=", 1)[1]
    if dataset is None:
        raise ValueError("Dataset name is not provided.")
    out: List[str] = []
    for arg in argv:
        if arg.startswith("@"):
            name = arg[1:]
            base = os.path.join(config_dir, "shortcut", name, "base.yaml")
            if os.path.exists(base):
                out += _yaml_to_cli(base)
                per_dataset = os.path.join(config_dir, "shortcut", name, f"{dataset}.yaml")
                if os.path.exists(per_dataset):
                    out += _yaml_to_cli(per_dataset)
            else:
                single = os.path.join(config_dir, "shortcut", f"{name}.yaml")
                if not os.path.exists(single):
                    raise ValueError(f"Shortcut @{name} not found.")
                out += _yaml_to_cli(single)
        elif arg.startswith("algorithm/backbone="):
            out += ["algorithm.backbone=null"] + _yaml_to_cli(
                os.path.join(config_dir, "algorithm", "backbone", arg.split("=", 1)[1] + ".yaml"), "algorithm.backbone")
        elif arg.startswith("algorithm/vae="):
            out += ["algorithm.vae=null"] + _yaml_to_cli(
                os.path.join(config_dir, "algorithm", arg.split("=", 1)[1] + ".yaml"), "algorithm.vae")
        else:
            out.append(arg)
    return out


# ------------------------------------------------------------------------------------------ defaults-list composition
class _Composer:
    def __init__(self, config_dir: str, cli_choices: Dict[str, str]):
        self.dir = config_dir
        self.cli_choices = cli_choices            # group path -> choice given on the command line
        self.choices: Dict[str, Optional[str]] = {}

    def _read(self, rel: str) -> Optional[Tuple[dict, Optional[str]]]:
        path = os.path.join(self.dir, rel + ".yaml")
        if not os.path.exists(path):
            return None
        with open(path) as f:
            text = f.read()
        header = re.match(r"\s*#\s*@package\s+(\S+)", text)
        return (_load_yaml(text) or {}), (header.group(1) if header else None)

    def _choice(self, group: str, choice, overrides: Dict[str, str]):
        choice = self.cli_choices.get(group, overrides.get(group, choice))
        if isinstance(choice, str):      # `${dataset}_${experiment}`: interpolation over the choices made so far
            choice = _INTERP.sub(lambda m: str(self.choices.get(m.group(1), m.group(0))), choice)
        return choice

    def load(self, group: str, name: str, package: str, overrides: Dict[str, str], optional: bool = False) -> dict:
        """The composed content of `<group>/<name>.yaml`, placed at `package`."""
        got = self._read(os.path.join(group, name) if group else name)
        if got is None:
            if optional:
                return {}
            raise FileNotFoundError(f"config '{os.path.join(group, name)}.yaml' not found under {self.dir}")
        body, header = got
        if header == "_global_":
            package = ""
        elif header not in (None, "_group_"):
            package = header
        defaults = body.pop("defaults", None) or []
        overrides = dict(overrides)
        for d in defaults:               # this file's `override g: c` entries steer the groups its defaults pull in;
            if isinstance(d, dict):      # overrides coming from further out (or the command line) win
                for k, v in d.items():
                    if k.startswith("override "):
                        overrides.setdefault(self._norm(group, k[len("override "):].strip().split("@")[0]), v)
        result: dict = {}
        self_done = False
        for d in defaults:
            if d == "_self_":
                _merge(result, _at_package(body, package))
                self_done = True
            elif isinstance(d, str):
                _merge(result, self.load(group, d, package, overrides))
            elif isinstance(d, dict):
                for key, choice in d.items():
                    if key.startswith("override "):
                        continue
                    opt = key.startswith("optional ")
                    key = key[len("optional "):].strip() if opt else key
                    rel, _, pkg = key.partition("@")
                    sub_group = self._norm(group, rel)
                    # Hydra keys a default with an explicit package as `group@package`: a plain `group=choice` on the
                    # command line (or an `override group:`) does not touch it
                    choice = self._choice(f"{sub_group}@{pkg}" if pkg else sub_group, choice, overrides)
                    if not pkg:
                        self.choices[sub_group] = choice
                    if choice is None:
                        continue
                    sub_package = pkg if pkg else ".".join(p for p in (package, rel.replace("/", ".")) if p)
                    if pkg and package and not pkg.startswith("_global_"):
                        sub_package = f"{package}.{pkg}"
                    sub_package = sub_package.replace("_global_.", "").replace("_global_", "")
                    _merge(result, self.load(sub_group, str(choice), sub_package, overrides, optional=opt))
        if not self_done:
            _merge(result, _at_package(body, package))
        return result

    @staticmethod
    def _norm(group: str, rel: str) -> str:
        return os.path.normpath(os.path.join(group, rel)).replace(os.sep, "/").lstrip("./")


# ------------------------------------------------------------------------------------------ interpolation
def _resolve(tree: dict) -> dict:
    unresolvable = ("now:", "hydra:", "hydra.", "oc.", "env:")
    cache: Dict[str, Any] = {}

    def lookup(path: str, stack: Tuple[str, ...]):
        if path in stack:
            raise ValueError(f"interpolation cycle: {' -> '.join(stack + (path,))}")
        if path not in cache:
            cache[path] = walk(_get(tree, path), stack + (path,))
        return cache[path]

    def walk(node, stack: Tuple[str, ...] = ()):
        if isinstance(node, dict):
            return {k: walk(v, stack) for k, v in node.items()}
        if isinstance(node, list):
            return [walk(v, stack) for v in node]
        if not isinstance(node, str) or "${" not in node:
            return node
        whole = _INTERP.fullmatch(node)
        if whole and not whole.group(1).startswith(unresolvable):
            try:
                return copy.deepcopy(lookup(whole.group(1).strip(), stack))
            except KeyError:
                return node

        def sub(m):
            key = m.group(1).strip()
            if key.startswith(unresolvable):
                return m.group(0)
            try:
                return str(lookup(key, stack))
            except KeyError:
                return m.group(0)
        prev = None
        while prev != node and "${" in node:
            prev, node = node, _INTERP.sub(sub, node)
        return node

    return walk(tree)


# ------------------------------------------------------------------------------------------ entry point
def compose(config_dir: str, argv: Optional[List[str]] = None, config_name: str = "config", resolve: bool = True) -> dict:
    """The reference's `python -m main <argv>` configuration as a plain dict (see the module docstring)."""
    argv = unwrap_shortcuts(list(argv or []), config_dir, config_name)
    with open(os.path.join(config_dir, f"{config_name}.yaml")) as f:
        root_defaults = (_load_yaml(f.read()) or {}).get("defaults", [])
    groups = set()
    for d in root_defaults:
        if isinstance(d, dict):
            for k in d:
                groups.add(k.replace("optional ", "").replace("override ", "").split("@")[0].strip())
    choices, edits = {}, []
    for arg in argv:
        if arg.startswith("~"):
            edits.append(("del", arg[1:].split("=")[0], None))
            continue
        if "=" not in arg:
            raise ValueError(f"cannot parse override '{arg}' (expected key=value)")
        key, value = arg.split("=", 1)
        mode = None if key.startswith("++") else (False if key.startswith("+") else True)
        key = key.lstrip("+")
        if mode is True and (key in groups or os.path.isdir(os.path.join(config_dir, key))) and "." not in key:
            choices[key] = None if value in ("null", "") else value
        else:
            edits.append(("set", key, (mode, parse_value(value))))
    composer = _Composer(config_dir, choices)
    tree = composer.load("", config_name, "", {})
    for op, key, payload in edits:
        if op == "del":
            parent, _, leaf = key.rpartition(".")
            (_get(tree, parent) if parent else tree).pop(leaf, None)
        else:
            _set(tree, key, payload[1], payload[0])
    for group in ("experiment", "dataset", "algorithm"):           # main.py:51-57
        if isinstance(tree.get(group), dict) and composer.choices.get(group) is not None:
            tree[group]["_name"] = composer.choices[group]
    tree.pop("hydra", None)
    return _resolve(tree) if resolve else tree
