"""Stage the UNMODIFIED reference so that it travels to the GPU box (test infrastructure only).

    python -m oracle.stage_ref            # build container only; needs /root/reference

``/root/reference`` does not exist on the GPU box.  The reference is a Python script tree (no ``setup.py`` /
``pyproject.toml``, so the base contract's ``pip install --target`` has nothing to install); this script is that
install step done by hand: it copies the python files of ``codes/`` byte for byte into ``oracle/_ref/codes/``.
``oracle/_ref/`` is git-ignored (the reference's sources never enter this repository's history) but NOT
gpurun-ignored, so the directory ships with the snapshot like the built ``.so``.  A manifest with the SHA-256 of every
staged file is written next to it; ``oracle/ref_loader.py`` refuses a staged tree whose files no longer match.

Who may use ``oracle/_ref``: ``tests/`` (the eager reference on the B200 as the parity oracle at full sizes, the
config-5 caller ``custom_video_test.py``), ``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline`` leg.
Never the product path.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("STIF_REFERENCE_ROOT", "/root/reference")
KEEP_EXT = (".py", ".yml", ".yaml")


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def stage(source: str = SOURCE, dest: str = STAGED) -> dict:
    """Copy ``<source>/codes/**/*.{py,yml}`` to ``<dest>/codes/`` and write ``<dest>/MANIFEST.json``."""
    src_codes = os.path.join(source, "codes")
    if not os.path.isdir(src_codes):
        raise RuntimeError(f"no reference checkout under {source}")
    if os.path.isdir(dest):
        shutil.rmtree(dest)
    manifest = {}
    for dirpath, dirnames, filenames in os.walk(src_codes):
        dirnames[:] = sorted(d for d in dirnames if d not in ("__pycache__", ".ipynb_checkpoints"))
        for fn in sorted(filenames):
            if not fn.endswith(KEEP_EXT):
                continue
            rel = os.path.relpath(os.path.join(dirpath, fn), source)
            out = os.path.join(dest, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(os.path.join(source, rel), out)
            manifest[rel] = _sha(out)
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": source, "files": manifest}, f, indent=1, sort_keys=True)
    return manifest


def verify(dest: str = STAGED) -> bool:
    """True iff every staged file still has the hash recorded when it was copied."""
    mpath = os.path.join(dest, "MANIFEST.json")
    if not os.path.isfile(mpath):
        return False
    files = json.load(open(mpath))["files"]
    return bool(files) and all(os.path.isfile(os.path.join(dest, rel)) and _sha(os.path.join(dest, rel)) == h
                               for rel, h in files.items())


if __name__ == "__main__":
    m = stage()
    print(f"staged {len(m)} reference files into {STAGED}", file=sys.stderr)
