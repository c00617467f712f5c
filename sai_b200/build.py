"""Builds ``sai_b200/lib/libsai_b200.so`` (the C-ABI library) in-tree with nvcc
for sm_100a.  No torch involvement: the library has no torch types."""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIB = ROOT / "lib" / "libsai_b200.so"
INCLUDE = ROOT.parent / "include"

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-lineinfo",
    "-std=c++17",
    "--fmad=false",  # float64 decisions must match numpy: never contract mul+add
    "-Xcompiler",
    "-fPIC,-O3,-pthread",
    "-cudart",
    "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; the sai_b200 CUDA library cannot be built")


def sources() -> list[Path]:
    """CUDA sources (nvcc) and plain C++ sources (g++: the host packer's x86 vector paths)."""
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cpp"))


GXX_FLAGS = ["-O3", "-std=c++17", "-fPIC", "-pthread", "-Wall"]


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.cpp")) + list(INCLUDE.glob("*.h"))
    return any(d.stat().st_mtime > t for d in deps)


EXPERIMENTS_LIB = ROOT.parent / "tools" / "bin" / "libsai_b200_exp.so"


def build(force: bool = False, verbose: bool = False, experiments: bool = False) -> Path:
    """``experiments=True`` builds ``tools/bin/libsai_b200_exp.so`` with -DSAI_EXPERIMENTS: the
    genotype-pass and window-kernel variants that were measured slower (profiles/round1_notes.md)
    stay available to the A/B scripts under tools/ without shipping in the product library
    (select it with ``SAI_B200_LIB=tools/bin/libsai_b200_exp.so``)."""
    if experiments:
        return _compile(EXPERIMENTS_LIB, ROOT.parent / "tools" / "bin" / "obj", ["-DSAI_EXPERIMENTS"], verbose)
    if not force and not is_stale():
        return LIB
    return _compile(LIB, ROOT / "lib" / "obj", [], verbose)


def _compile(LIB: Path, obj_dir: Path, defines: list, verbose: bool) -> Path:
    LIB.parent.mkdir(parents=True, exist_ok=True)
    objs = []
    obj_dir.mkdir(parents=True, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    for src in sources():
        obj = obj_dir / (src.stem + ".o")
        if src.suffix == ".cpp":
            cmd = [os.environ.get("CXX", "g++"), *GXX_FLAGS, *defines, "-I", str(INCLUDE), "-c", str(src), "-o", str(obj)]
        else:
            cmd = [nvcc, *NVCC_FLAGS, *defines, "-I", str(INCLUDE), "-c", str(src), "-o", str(obj)]
        if verbose and src.suffix != ".cpp":
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out.decode(errors="replace"))
        if p.returncode != 0:
            raise RuntimeError(f"compiling {src.name} failed")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static",
           "-Xcompiler", "-fPIC,-pthread", "-o", str(LIB)] + [
        str(o) for o in objs
    ] + ["-lz"]  # zlib: BGZF blocks of .vcf.gz inputs (host ingest)
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, experiments="--experiments" in sys.argv))
