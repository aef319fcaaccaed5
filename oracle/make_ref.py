"""Recipe for ``oracle/_ref``: the UNMODIFIED reference modules of the hot path, staged so that they travel to the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY (same rule as ``cnf_oracle.py``: only ``tests/``, ``__graft_entry__`` and
``bench.py``'s CPU legs may use anything under ``oracle/``).

    python oracle/make_ref.py            # copies /root/reference/src/<hot modules>.py -> oracle/_ref/src/

The reference is pure Python (no build step), so "building" it is copying its source files byte for byte from where
they lie under ``/root/reference`` into ``oracle/_ref/src`` -- a directory that is git-ignored (it never enters the
history) but not gpurun-ignored (it travels with the snapshot like our own built ``.so``).  ``oracle/ref_driver.py``
imports the staged modules (with stub modules for the absent visualisation dependencies) and drives them exactly as
``train.py:88-89,123-182`` does; ``bench.py --impl reference`` and ``cpu_baseline`` time that.  A SHA-256 manifest is
written next to the copies so that a run can state which bytes it timed.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CNF_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref", "src")
# hot-path modules + what their import statements pull in (trainer -> vis; scene_cateogries -> trainer, utils)
FILES = ["render_rays.py", "embedding.py", "model.py", "loss.py", "trainer.py", "scene_cateogries.py", "utils.py", "vis.py"]


def stage(verbose=True):
    """Copy the reference's hot-path modules into oracle/_ref/src.  Returns the destination or None if there is no reference."""
    src_dir = os.path.join(REF, "src")
    if not os.path.isdir(src_dir):
        return DST if os.path.exists(os.path.join(DST, "MANIFEST.json")) else None
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for f in FILES:
        s, d = os.path.join(src_dir, f), os.path.join(DST, f)
        shutil.copyfile(s, d)
        manifest[f] = hashlib.sha256(open(d, "rb").read()).hexdigest()
    json.dump({"source": src_dir, "sha256": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print("staged %d reference modules -> %s" % (len(FILES), DST))
    return DST


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
