"""Recipe: put the reference's OWN implementation of the path where the benchmark arms can run it.

The reference (WujiangXu/AMID) is six pure-Python files with no build system, so "building" it is a byte copy:
`python oracle/build_ref.py` (also called by __graft_entry__.build()) copies the files named below from
/root/reference into oracle/_ref/ and records their SHA-256 digests in oracle/_ref/MANIFEST.json.  oracle/_ref/
is git-ignored (no reference source ever enters this repository's history) but it is NOT gpurun-ignored, so the
copy travels to the GPU box with the working tree, where /root/reference does not exist.  Nothing is modified:
the two compatibility shims of SURVEY.md section 8c are applied at import time by oracle/ref_loader.py.

Test / benchmark infrastructure only -- nothing under amid_b200/ may import from oracle/.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
CODE = ["model_seq.py", "dataset_seq.py", "utils.py", "train_sr.py", "train_sr_dr.py", "run.sh"]
# the cloth_sport files BASELINE configs 1-2 name (run.sh defaults / train_sr_dr.py:636-642); 3.2 MB
DATA = ["amazon_dataset/cloth_sport_train75.csv", "amazon_dataset/cloth_sport_train75_DR.csv",
        "amazon_dataset/cloth_sport_test.csv"]


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        for blk in iter(lambda: fh.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def available() -> bool:
    return os.path.exists(os.path.join(DST, "MANIFEST.json"))


def build(verbose: bool = False) -> str:
    """Copy the reference into oracle/_ref (when /root/reference is present); returns a one-line status."""
    if not os.path.isdir(SRC):
        return "oracle/_ref: present (prebuilt)" if available() else "oracle/_ref: absent (no /root/reference on this machine)"
    manifest = {}
    for rel in CODE + DATA:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and os.path.getsize(dst) == os.path.getsize(src) and _sha(dst) == _sha(src)):
            shutil.copyfile(src, dst)
            os.chmod(dst, 0o644)
        manifest[rel] = _sha(dst)
        if verbose:
            print(f"{rel}: {manifest[rel][:16]}")
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": "WujiangXu/AMID (read-only copy at /root/reference)", "files": manifest}, fh, indent=1)
    return f"oracle/_ref: {len(manifest)} reference files copied"


if __name__ == "__main__":
    print(build(verbose=True))
