"""Recipe for oracle/_ref/: the reference's OWN hot-path sources, taken from where they lie under /root/reference.

TEST / BASELINE INFRASTRUCTURE ONLY (never imported by ditto_tts_b200/).  The reference is pure Python, so "building" it is
placing the two files of the path -- src/components/DiT.py and src/model/DiTTO.py, byte for byte -- under oracle/_ref/src/
(git-ignored: no reference source enters the history; NOT gpurun-ignored: the tree travels to the GPU box like a built .so).
oracle/ref_loader.py imports them there with the NAC constructor stubbed and bench.py times them as
the `--impl reference` arm / `cpu_baseline` (kind "reference").

    python oracle/build_ref.py        # needs /root/reference (the build container); prints the sha256 of what it placed
"""
from __future__ import annotations

import hashlib
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference/src"
DEST = os.path.join(HERE, "_ref", "src")
FILES = ["components/DiT.py", "model/DiTTO.py"]


def build(verbose: bool = True) -> bool:
    """Returns True when oracle/_ref/ is complete (freshly placed, or already there and /root/reference absent)."""
    have_src = all(os.path.exists(os.path.join(REF_ROOT, f)) for f in FILES)
    if not have_src:
        return all(os.path.exists(os.path.join(DEST, f)) for f in FILES)
    manifest = []
    for f in FILES:
        dst = os.path.join(DEST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF_ROOT, f), dst)
        with open(dst, "rb") as fh:
            manifest.append(f"{hashlib.sha256(fh.read()).hexdigest()}  src/{f}")
    with open(os.path.join(HERE, "_ref", "MANIFEST.sha256"), "w") as fh:
        fh.write("\n".join(manifest) + "\n")
    if verbose:
        print("oracle/_ref:", *manifest, sep="\n  ")
    return True


if __name__ == "__main__":
    ok = build()
    raise SystemExit(0 if ok else 1)
