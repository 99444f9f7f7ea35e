"""Stage the UNMODIFIED reference sources of the hot path under oracle/_ref/ (TEST / BENCH INFRASTRUCTURE).

    python -m oracle.make_ref

The reference is pure Python: there is nothing to compile, so "building oracle/_ref" means copying the
files the path is made of, byte for byte, from /root/reference (read-only, build container only) into
oracle/_ref/, which is git-ignored (never part of the history) but travels to the GPU box with the
snapshot like a built .so.  ``oracle/reference_shim.py`` then imports them there exactly as it does
from /root/reference, so ``bench.py --impl reference`` and ``cpu_baseline`` can time the reference's own
code on the GPU box's host cores (cpu_baseline.kind = "reference").  A manifest with the SHA-256 of every
staged file is written next to them; ``staged_ok()`` re-verifies it before use.
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = "/root/reference"
FILES = ["mfcc.py", "utils.py", "config.py",
         "dataset/file_processing.py", "dataset/sph.py", "dataset/stm_parser.py", "dataset/file_index.py",
         "dataset/utils.py", "realtime_analysis/analyser.py", "realtime_analysis/sklearn_analyser.py"]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage():
    """Copy the files (no-op without /root/reference).  Returns True when oracle/_ref is usable."""
    if not os.path.isfile(os.path.join(SRC, "mfcc.py")):
        return staged_ok()
    manifest = {}
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1, sort_keys=True)
    return True


def staged_ok():
    man = os.path.join(DST, "MANIFEST.json")
    if not os.path.isfile(man):
        return False
    with open(man) as f:
        files = json.load(f)["files"]
    return all(os.path.isfile(os.path.join(DST, rel)) and _sha(os.path.join(DST, rel)) == h
               for rel, h in files.items()) and set(files) == set(FILES)


if __name__ == "__main__":
    print("oracle/_ref staged:", stage())
