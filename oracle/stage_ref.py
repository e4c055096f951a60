"""Stage the UNMODIFIED reference for the reference arm  --  TEST / MEASUREMENT INFRASTRUCTURE ONLY.

    python oracle/stage_ref.py          (authoring container; needs /root/reference)

The reference is a pure-Python script tree (no setup.py, nothing to compile), so "building" it means placing the files
of the hot path where ``bench.py --impl reference`` and the tests can import them on the GPU box, which has no
/root/reference: the upstream classes ``legacy_models/*.py`` under the package name ``models`` they import each other
by next to ``utils/``.  They are copied byte for byte from where they lie into ``oracle/_ref/``, which is
listed in .gitignore (it never enters the history) but not in .gpurunignore (it travels with the snapshot, like the
built .so).  ``oracle/_ref/MANIFEST.json`` records the source path and SHA-256 of every file, so a reader can check the
copy is unmodified.  Nothing in the product path imports it.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("XNV2_REFERENCE_ROOT", "/root/reference")
# (source directory, package name under oracle/_ref)
TREES = [("legacy_models", "models"), ("utils", "utils")]
# models/ (the fork's refactored classes) is staged too, under its own name, for the Captioner call style
EXTRA = [("models", "models_refactored")]


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def stage(force: bool = False) -> str:
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "legacy_models")):
        if os.path.isdir(os.path.join(DEST, "models")):
            return DEST                      # already staged (GPU box)
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT} and nothing staged at {DEST}")
    if os.path.isdir(DEST) and not force:
        try:
            with open(os.path.join(DEST, "MANIFEST.json")) as f:
                man = json.load(f)
            if all(os.path.exists(os.path.join(DEST, e["staged"])) and _sha(os.path.join(REFERENCE_ROOT, e["source"])) == e["sha256"]
                   for e in man["files"]):
                return DEST
        except Exception:
            pass
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    files = []
    for src, pkg in TREES + EXTRA:
        for root, _, names in os.walk(os.path.join(REFERENCE_ROOT, src)):
            for n in sorted(names):
                if not n.endswith(".py"):
                    continue
                sp = os.path.join(root, n)
                rel = os.path.relpath(sp, os.path.join(REFERENCE_ROOT, src))
                dp = os.path.join(DEST, pkg, rel)
                os.makedirs(os.path.dirname(dp), exist_ok=True)
                shutil.copyfile(sp, dp)
                files.append({"source": os.path.relpath(sp, REFERENCE_ROOT), "staged": os.path.relpath(dp, DEST), "sha256": _sha(sp)})
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"reference_root": REFERENCE_ROOT, "files": files}, f, indent=1)
    return DEST


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
