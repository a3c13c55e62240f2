"""Recipe for oracle/_ref: the reference's OWN implementation of the hot path, made runnable on the GPU box's host cores.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference is pure Python: its hot-path arithmetic lives in nine files under
`/root/reference/src/modules` plus `src/config/models.yaml`; the per-frame loop that calls them is
`src/can_swap_pipeline_e2e.py` (+ the helpers it imports from `src/utils`, `src/config`).  `/root/reference` does not exist on the GPU box, so this
script copies exactly those files, unmodified, from where they lie into `oracle/_ref/` (git-ignored -- reference sources
never enter the repository history -- but shipped to the box with the snapshot, like the built .so files).  It is run by
`__graft_entry__.build()` whenever `/root/reference` is present.

Consumers: `bench.py --impl reference` and bench.py's `cpu_baseline` leg (kind "reference" when oracle/_ref exists, else the
oracle port), and tests/test_ref_bundle.py (the bundle equals the oracle and the live reference).  Nothing under
`canonswap_b200/` may import it.

    python oracle/make_ref.py            # (re)create oracle/_ref
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CANONSWAP_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = [
    "src/modules/__init__.py",
    "src/modules/util.py",
    "src/modules/appearance_feature_extractor.py",
    "src/modules/dense_motion.py",
    "src/modules/warping_network.py",
    "src/modules/spade_generator.py",
    "src/modules/adaptive_modulate.py",
    "src/modules/convnextv2.py",
    "src/modules/motion_extractor.py",
    "src/modules/stitching_retargeting_network.py",     # imported by src/utils/helper.py
    "src/config/models.yaml",
    "src/config/__init__.py",
    "src/config/base_config.py",
    "src/config/argument_config.py",
    "src/config/inference_config.py",
    "src/config/crop_config.py",
    "src/utils/camera.py",
    "src/utils/helper.py",
    "src/utils/crop.py",
    "src/utils/rprint.py",
    # the caller of the path: tests/ref_pipeline_harness.py executes its LOOP C text (:223-283) and make_motion_template
    # (:101-135) against the drop-in can_swapper, with the I/O imports stubbed
    "src/can_swap_pipeline_e2e.py",
    "src/can_swap_pipeline_v2i.py",
]


def make(verbose: bool = True) -> str | None:
    if not os.path.isdir(os.path.join(REF, "src", "modules")):
        if verbose:
            print(f"make_ref: {REF} not present; keeping whatever oracle/_ref already holds")
        return DST if os.path.isdir(DST) else None
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for rel in FILES:
        src = os.path.join(REF, rel)
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    for pkg in ("src", "src/utils", "src/config"):              # namespace markers the copied files expect
        init = os.path.join(DST, pkg, "__init__.py")
        if not os.path.exists(init):
            open(init, "w").close()
    json.dump({"reference": REF, "sha256": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print(f"make_ref: {len(FILES)} reference files -> {DST}")
    return DST


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
