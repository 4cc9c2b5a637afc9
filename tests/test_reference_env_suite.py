"""The reference's OWN environment test file (tests/test_env.py, unmodified) executed against the facade
``mapf_gnn_b200.GraphEnv``: the module name it imports (``grid.env_graph_gridv1``) is bound to a shim that exports the
facade's ``GraphEnv / create_goals / create_obstacles``.  The file comes from ``oracle/_ref/tests/`` (staged, digest
checked, travels to the GPU box) or from ``/root/reference`` (build container).  Every ``test_*`` of the file runs,
including its timing bounds (step < 5 ms, collision check < 1 ms)."""
import importlib.util
import inspect
import os
import sys
import types

import pytest

from oracle import ref_loader

pytestmark = pytest.mark.gpu


def _load_reference_tests():
    root = ref_loader.root()
    if root is None:
        pytest.skip("reference tests neither staged under oracle/_ref nor present at /root/reference")
    path = os.path.join(root, "tests", "test_env.py")
    if not os.path.exists(path):
        pytest.skip(f"{path} is missing")
    import mapf_gnn_b200 as m
    shim = types.ModuleType("grid.env_graph_gridv1")
    shim.GraphEnv, shim.create_goals, shim.create_obstacles = m.GraphEnv, m.create_goals, m.create_obstacles
    pkg = types.ModuleType("grid")
    pkg.env_graph_gridv1 = shim
    saved = {k: sys.modules.get(k) for k in ("grid", "grid.env_graph_gridv1")}
    sys.modules["grid"], sys.modules["grid.env_graph_gridv1"] = pkg, shim
    try:
        spec = importlib.util.spec_from_file_location("reference_test_env", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    assert mod.GraphEnv is m.GraphEnv
    return mod


def _cases():
    """(class name or None, function name) of every test in the reference file; resolved at collection time from the
    staged copy when it exists (so the ids are visible), else one placeholder that skips."""
    root = ref_loader.root()
    path = None if root is None else os.path.join(root, "tests", "test_env.py")
    if path is None or not os.path.exists(path):
        return [("-", "-")]
    import ast
    tree = ast.parse(open(path).read())
    out = []
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name.startswith("Test"):
            out += [(node.name, f.name) for f in node.body if isinstance(f, ast.FunctionDef) and f.name.startswith("test_")]
        elif isinstance(node, ast.FunctionDef) and node.name.startswith("test_"):
            out.append((None, node.name))
    return out


@pytest.fixture(scope="module")
def reference_tests():
    return _load_reference_tests()


@pytest.mark.parametrize("cls,name", _cases())
def test_reference_env_test(reference_tests, cls, name):
    if name == "-":
        pytest.skip("reference test file not available")
    if cls is None:
        getattr(reference_tests, name)()
        return
    inst = getattr(reference_tests, cls)()
    if hasattr(inst, "setup_method"):
        inst.setup_method()
    getattr(inst, name)()
