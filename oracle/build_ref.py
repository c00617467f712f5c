#!/usr/bin/env python
"""TEST / BENCH INFRASTRUCTURE -- installs the UNMODIFIED reference into oracle/_ref/.

The reference (xin-huang/sai) is pure Python: there is nothing to compile, "building" it is a
`pip install --no-deps --target oracle/_ref` of /root/reference (from a scratch copy, because
setuptools writes egg-info into the source tree and /root/reference is read-only).  oracle/_ref/
is git-ignored (no reference source ever enters the history) but NOT gpurun-ignored, so the
installed package travels to the GPU box with the snapshot -- /root/reference does not exist
there.  Its third-party dependencies scikit-allel, pysam and natsort are not installed in this
image; `oracle/ref_driver.py` stubs them at import (none of them is executed on the in-memory
scoring path: `allel` is only imported by sai/stats/stat_utils.py:22, `pysam` only by
ChunkGenerator.__init__, `natsort` only by `outlier`).

    python oracle/build_ref.py            # no-op when oracle/_ref is already there
    python oracle/build_ref.py --force

Called by `__graft_entry__.build()` when /root/reference is present.  Only tests/, smoke() and
bench.py's CPU legs (`--impl reference`, `cpu_baseline`) ever import oracle/_ref.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
REFERENCE = os.environ.get("SAI_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(DEST, "sai", "stats", "stat_utils.py"))


def build(force: bool = False) -> str | None:
    """Returns the install directory, or None when the reference is not mounted and nothing was
    installed earlier."""
    if available() and not force:
        return DEST
    if not os.path.isdir(os.path.join(REFERENCE, "sai")):
        return DEST if available() else None
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REFERENCE, src, ignore=shutil.ignore_patterns(".git", "docs", "examples", "tests"))
        if os.path.isdir(DEST):
            shutil.rmtree(DEST)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--ignore-requires-python",  # the reference pins python ==3.9.19; this image has 3.12
               "--find-links", "/opt/wheelhouse", "--target", DEST, src]
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if p.returncode != 0 or not available():
            sys.stderr.write(p.stdout)
            raise RuntimeError("pip install of the reference into oracle/_ref failed")
    with open(os.path.join(DEST, "INSTALLED_FROM"), "w") as f:
        f.write(f"{REFERENCE} (pip install --no-deps --target; unmodified)\n")
    return DEST


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
