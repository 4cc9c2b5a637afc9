"""Test infrastructure -- imports the unmodified reference classes (``GraphEnv`` and the three ``Network`` modules).

Search order: ``oracle/_ref/`` (staged by ``oracle/stage_ref.py``; digest-checked against its manifest), then
``/root/reference`` (build container).  ``available()`` is False on a box that has neither; callers then fall back to
the numpy / torch-CPU port in ``oracle/env_oracle.py`` / ``oracle/model_oracle.py`` and say so
(``cpu_baseline.kind == "port"``).

Import shims: ``grid/env_graph_gridv1.py:3-6`` imports matplotlib and gymnasium at module import and uses
them only in ``__init__`` / ``render``; inert stub modules are registered when the real packages are missing.
``models/framework_gnn_message.py`` / ``framework_baseline.py`` import ``networks.*`` relative to ``models/``.
Only ``tests/``, ``__graft_entry__`` and ``bench.py``'s CPU legs may import this module.
"""
from __future__ import annotations

import hashlib
import importlib
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")
_cache = {}


def _staged_ok():
    man = os.path.join(STAGED, "MANIFEST.json")
    if not os.path.exists(man):
        return False
    with open(man) as fh:
        digests = json.load(fh)["sha256"]
    for rel, want in digests.items():
        p = os.path.join(STAGED, rel)
        if not os.path.exists(p):
            return False
        with open(p, "rb") as fh:
            if hashlib.sha256(fh.read()).hexdigest() != want:
                raise RuntimeError(f"oracle/_ref/{rel} differs from the staged reference (modified?)")
    return True


def root():
    """Directory holding the reference tree, or None."""
    if "root" not in _cache:
        r = None
        if _staged_ok():
            r = STAGED
        else:
            cand = os.environ.get("MAPF_REFERENCE", "/root/reference")
            if os.path.isdir(os.path.join(cand, "grid")):
                r = cand
        _cache["root"] = r
    return _cache["root"]


def available() -> bool:
    return root() is not None


def _install_shims():
    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Anything()

        def __getattr__(self, name):
            return _Anything()

    def stub(name, **attrs):
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)

        def _getattr(attr):
            if attr.startswith("__"):
                raise AttributeError(attr)
            return _Anything()

        mod.__getattr__ = _getattr                          # type: ignore[attr-defined]
        sys.modules[name] = mod
        return mod

    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl = stub("matplotlib")
        mpl.pyplot = stub("matplotlib.pyplot")
        mpl.cm = stub("matplotlib.cm")
        mpl.colors = stub("matplotlib.colors")
    try:
        import gymnasium  # noqa: F401
    except Exception:
        gym = stub("gymnasium", Env=object)
        gym.spaces = stub("gymnasium.spaces")


def load():
    """Returns a namespace with GraphEnv, create_goals, create_obstacles, GNNNetwork, MessageNetwork, BaselineNetwork
    (the reference's three ``Network`` classes) and ``root``."""
    if "ns" in _cache:
        return _cache["ns"]
    r = root()
    if r is None:
        raise RuntimeError("the reference is neither staged under oracle/_ref nor present at /root/reference")
    _install_shims()
    for p in (os.path.join(r, "models"), r):
        if p not in sys.path:
            sys.path.insert(0, p)
    env_mod = importlib.import_module("grid.env_graph_gridv1")
    fg = importlib.import_module("models.framework_gnn")
    fm = importlib.import_module("framework_gnn_message")
    fb = importlib.import_module("framework_baseline")
    ns = types.SimpleNamespace(root=r, GraphEnv=env_mod.GraphEnv, create_goals=env_mod.create_goals,
                               create_obstacles=env_mod.create_obstacles, GNNNetwork=fg.Network,
                               MessageNetwork=fm.Network, BaselineNetwork=fb.Network, env_module=env_mod)
    _cache["ns"] = ns
    return ns


def build_network(config):
    """evaluate.py:163-173: the Network class a config selects."""
    ns = load()
    if config.get("net_type") == "baseline":
        return ns.BaselineNetwork(config)
    if config.get("msg_type") == "message":
        return ns.MessageNetwork(config)
    return ns.GNNNetwork(config)
