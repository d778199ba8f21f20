"""Stage the reference's python files of the hot path under baseline/_ref/ (git-ignored; travels to the GPU box).

    python -m oracle.stage_reference

TEST INFRASTRUCTURE.  /root/reference does not exist on the GPU box, and the drop-in claim -- the UNMODIFIED
callers run on `import pointops` from this repo -- can only be tested there if the callers themselves are
present.  BASELINE.md 3.2 reserves baseline/_ref/ for the reference install; the repository has no package
metadata to pip-install, so the install is this copy of exactly the files oracle/ref_glue.py loads, byte for
byte, at their original relative paths.  Nothing under baseline/_ref/ is tracked by git.
"""
import hashlib
import json
import os
import shutil
import sys

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("POINTCLOUDPDF_REFERENCE", "/root/reference")
DST = os.path.join(REPO_ROOT, "baseline", "_ref")
FILES = [
    "libs/pointops/functions/__init__.py",
    "libs/pointops/functions/aggregation.py",
    "libs/pointops/functions/attention.py",
    "libs/pointops/functions/grouping.py",
    "libs/pointops/functions/interpolation.py",
    "libs/pointops/functions/query.py",
    "libs/pointops/functions/sampling.py",
    "libs/pointops/functions/subtraction.py",
    "libs/pointops/functions/utils.py",
    "pointcept/models/point_transformer/point_transformer_seg.py",
    "pointcept/models/point_transformer/utils.py",
    "pointcept/recognizers/max_probability/max_probability_v1m1_base.py",
    "pointcept/recognizers/recognizer_model/pt_v1.py",
]


def stage() -> dict:
    if not os.path.isdir(SRC):
        raise SystemExit(f"{SRC} not present: nothing to stage (the GPU box uses the staged copy)")
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    json.dump(manifest, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    return manifest


if __name__ == "__main__":
    m = stage()
    print(f"staged {len(m)} files under {DST}")
