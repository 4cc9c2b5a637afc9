"""Test infrastructure -- stages the UNMODIFIED reference sources of the hot path into ``oracle/_ref/``.

    python oracle/stage_ref.py            # needs /root/reference (build container); idempotent

``oracle/_ref/`` is git-ignored (reference sources never enter this repository's history) but not gpurun-ignored, so
the byte-identical files travel to the GPU box, where ``bench.py --impl reference`` and ``cpu_baseline`` time the
reference's own classes (``cpu_baseline.kind == "reference"``) and ``tests/test_reference_env_suite.py`` runs the
reference's ``tests/test_env.py`` against the facade.  A manifest of SHA-256 digests is written next to the copies;
``oracle/ref_loader.py`` refuses files whose digest no longer matches (i.e. a modified reference).
Nothing in the product path (``mapf_gnn_b200/``) imports from here.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

REF = os.environ.get("MAPF_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")

# the path of SURVEY.md 8a/8b: the environment, the three Network modules with their layers and init, the callers'
# protocol files that are imported for constants only, and the reference's own environment test file
FILES = [
    "grid/__init__.py", "grid/env_graph_gridv1.py",
    "models/__init__.py", "models/framework_gnn.py", "models/framework_gnn_message.py", "models/framework_baseline.py",
    "models/networks/__init__.py", "models/networks/gnn.py", "models/networks/utils_weights.py",
    "data_loader.py",
    "tests/test_env.py",
    # the expert: the search, the instance generator / schedule parser and the reference's own tests
    "cbs/__init__.py", "cbs/cbs.py", "cbs/a_star.py",
    "data_generation/__init__.py", "data_generation/dataset_gen.py", "data_generation/trajectory_parser.py",
    "tests/test_cbs.py", "tests/test_cbs_performance.py", "tests/test_data_pipeline.py",
]


def stage(verbose=True) -> bool:
    if not os.path.isdir(REF):
        if verbose:
            print(f"{REF} is absent: nothing staged (existing oracle/_ref is kept)")
        return os.path.isdir(DST)
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as fh:
            manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REF, "sha256": manifest}, fh, indent=1, sort_keys=True)
    if verbose:
        print(f"staged {len(manifest)} reference files into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
