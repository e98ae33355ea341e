#!/usr/bin/env python
"""Recipe: stage the UNMODIFIED reference under oracle/_ref/ so that it travels to the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY (never imported by lithographysimulator_b200).

The reference is four flat Python modules with no packaging (nothing for pip to install), and
/root/reference is not mounted on the GPU box.  This script copies the four modules byte for byte from
where they lie (LITHO_REFERENCE or /root/reference) into oracle/_ref/ -- git-ignored, so no reference
source enters the history, but shipped by gpurun like every other built artefact -- and records their
sha256 in oracle/_ref/MANIFEST.json.  __graft_entry__.build() runs it whenever the reference is present.

    python oracle/build_ref.py            # stage (idempotent)
    python oracle/build_ref.py --check    # exit 0 iff oracle/_ref matches the manifest

Users: oracle/ref_runner.py (bench.py --impl reference, bench.py's cpu_baseline / library_baseline legs).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
MODULES = ("imageformation.py", "mask.py", "pupil.py", "lightsource.py")


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def staged() -> bool:
    """True when oracle/_ref holds the four modules and they match the manifest written at staging time."""
    man = os.path.join(DST, "MANIFEST.json")
    if not os.path.exists(man):
        return False
    try:
        m = json.load(open(man))
        return all(_sha(os.path.join(DST, n)) == m["sha256"][n] for n in MODULES)
    except Exception:
        return False


def stage(src: str | None = None) -> bool:
    src = src or os.environ.get("LITHO_REFERENCE", "/root/reference")
    if not all(os.path.exists(os.path.join(src, n)) for n in MODULES):
        return False
    os.makedirs(DST, exist_ok=True)
    sha = {}
    for n in MODULES:
        shutil.copyfile(os.path.join(src, n), os.path.join(DST, n))
        os.chmod(os.path.join(DST, n), 0o644)
        sha[n] = _sha(os.path.join(DST, n))
    json.dump({"source": src, "sha256": sha, "note": "verbatim copies; never edited, never committed"},
              open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    return True


if __name__ == "__main__":
    if "--check" in sys.argv:
        sys.exit(0 if staged() else 1)
    ok = stage()
    print("oracle/_ref staged" if ok else "reference not found: oracle/_ref not staged")
    sys.exit(0 if ok else 1)
