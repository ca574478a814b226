"""TEST INFRASTRUCTURE ONLY -- stage the UNMODIFIED reference where the GPU box can see it.

    python oracle/make_ref.py            (also called by __graft_entry__.build() when /root/reference is mounted)

The reference (yale-sys/prompt-cache) is pure Python; "/root/reference" exists only in the build container.  This
recipe copies its package `promptcache/` and the example schemas `examples/` byte for byte into `oracle/_ref/`
(git-ignored, NOT gpurun-ignored, so it travels with the snapshot like the built .so) and records a manifest with the
SHA-256 of every file.  Nothing is edited: the shims that make it importable under transformers 5 / without lxml live
in `oracle/ref_shim.py` and `oracle/_shims/`.  On the GPU box `oracle/ref_shim.py` finds `oracle/_ref/` and
  * `bench.py --impl reference` runs the reference's own CacheEngine.process -> GenerationEngine.generate on the host
    cores (`cpu_baseline.kind == "reference"`),
  * the `-m gpu` tests run the reference's own attention modules / models in fp16 ON THE B200 beside the CUDA path.
The product (`promptcache_b200/`) never imports anything from here.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("PC_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
WHAT = ("promptcache", "examples")


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def make_ref(verbose: bool = False) -> str:
    """Copy the reference package into oracle/_ref/ (idempotent).  Returns the destination, or '' when the source is
    absent (GPU box: the prebuilt copy that travelled with the snapshot is used as is)."""
    if not os.path.isdir(os.path.join(SRC, "promptcache")):
        return DST if os.path.isdir(os.path.join(DST, "promptcache")) else ""
    manifest = {}
    for sub in WHAT:
        src = os.path.join(SRC, sub)
        if not os.path.isdir(src):
            continue
        for root, dirs, files in os.walk(src):
            dirs[:] = [d for d in dirs if d != "__pycache__"]
            for f in files:
                if f.endswith((".pyc", ".pyo")):
                    continue
                s = os.path.join(root, f)
                rel = os.path.relpath(s, SRC)
                d = os.path.join(DST, rel)
                digest = _sha(s)
                manifest[rel] = digest
                if os.path.exists(d) and _sha(d) == digest:
                    continue
                os.makedirs(os.path.dirname(d), exist_ok=True)
                shutil.copyfile(s, d)
                if verbose:
                    print("copied", rel)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    return DST


def verify() -> bool:
    """True when every file of oracle/_ref matches the manifest written at copy time (unmodified reference)."""
    mf = os.path.join(DST, "MANIFEST.json")
    if not os.path.exists(mf):
        return False
    with open(mf) as fh:
        files = json.load(fh)["files"]
    return all(os.path.exists(os.path.join(DST, rel)) and _sha(os.path.join(DST, rel)) == d for rel, d in files.items())


if __name__ == "__main__":
    out = make_ref(verbose="-v" in sys.argv)
    print(out or "reference source not found and no staged copy present")
