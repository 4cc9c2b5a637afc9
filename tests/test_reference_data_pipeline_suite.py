"""The reference's OWN data-pipeline test file (tests/test_data_pipeline.py, unmodified) executed against
``mapf_gnn_b200.cbs``: ``data_generation.dataset_gen`` / ``data_generation.trajectory_parser`` are bound to shims
exporting this package's ``gen_input / data_gen / create_solutions / parse_trajectories / ...``.  One test of the file
inspects the reference's own source text (path separators in data_generation/*.py) and is not about behaviour: skipped.
Host code: no GPU needed."""
import ast
import importlib.util
import os
import sys
import types

import pytest

from oracle import ref_loader

SOURCE_TEXT_TESTS = {"test_no_hardcoded_separators"}


def _path():
    root = ref_loader.root()
    if root is None:
        return None
    p = os.path.join(root, "tests", "test_data_pipeline.py")
    return p if os.path.exists(p) else None


def _load():
    path = _path()
    if path is None:
        pytest.skip("reference tests neither staged under oracle/_ref nor present at /root/reference")
    import mapf_gnn_b200.cbs as mine
    pkg = types.ModuleType("data_generation")
    gen = types.ModuleType("data_generation.dataset_gen")
    par = types.ModuleType("data_generation.trajectory_parser")
    gen.gen_input, gen.data_gen, gen.create_solutions = mine.gen_input, mine.data_gen, mine.create_solutions
    par.get_longest_path, par.parse_trajectories = mine.get_longest_path, mine.parse_trajectories
    par.parse_dataset_trajectories = mine.parse_dataset_trajectories
    pkg.dataset_gen, pkg.trajectory_parser = gen, par
    names = ("data_generation", "data_generation.dataset_gen", "data_generation.trajectory_parser")
    saved = {k: sys.modules.get(k) for k in names}
    sys.modules.update(dict(zip(names, (pkg, gen, par))))
    try:
        spec = importlib.util.spec_from_file_location("reference_test_data_pipeline", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    assert mod.gen_input is mine.gen_input
    return mod


def _cases():
    path = _path()
    if path is None:
        return [("-", "-")]
    out = []
    for node in ast.parse(open(path).read()).body:
        if isinstance(node, ast.ClassDef) and node.name.startswith("Test"):
            out += [(node.name, f.name) for f in node.body if isinstance(f, ast.FunctionDef) and f.name.startswith("test_")]
        elif isinstance(node, ast.FunctionDef) and node.name.startswith("test_"):
            out.append((None, node.name))
    return out


@pytest.fixture(scope="module")
def reference_tests():
    return _load()


@pytest.mark.parametrize("cls,name", _cases())
def test_reference_data_pipeline_test(reference_tests, cls, name):
    if name == "-":
        pytest.skip("reference test file not available")
    if name in SOURCE_TEXT_TESTS:
        pytest.skip("reads the reference's own source files, not behaviour")
    if cls is None:
        getattr(reference_tests, name)()
        return
    inst = getattr(reference_tests, cls)()
    if hasattr(inst, "setup_method"):
        inst.setup_method()
    getattr(inst, name)()
