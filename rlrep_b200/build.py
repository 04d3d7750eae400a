"""Builds librlrep_b200.so (the C-ABI library) in-tree with nvcc for sm_100a.

    python -m rlrep_b200.build [--force] [--verbose]

The .so lands next to this file so it travels with the repo snapshot to the GPU box; it is git-ignored.
cudart is linked statically: the library has no dependency on torch or on a system libcudart.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = HERE / "_build"
LIB = HERE / "librlrep_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-I", str(HERE.parent / "include"),
] + (["-DRLREP_GEMM_TRACE"] if os.environ.get("RLREP_GEMM_TRACE") else [])


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _newer(target: Path, deps) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(d).stat().st_mtime <= t for d in deps)


def _deps(src: Path, seen=None):
    """Transitive closure of the quoted #includes of `src` inside csrc/ and include/ (so touching agent.cuh does not
    recompile the GEMM units)."""
    seen = set() if seen is None else seen
    for line in src.read_text().splitlines():
        line = line.strip()
        if line.startswith('#include "'):
            name = line.split('"')[1]
            for base in (CSRC, HERE.parent / "include"):
                h = base / name
                if h.exists() and h not in seen:
                    seen.add(h)
                    _deps(h, seen)
    return seen


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    OBJ.mkdir(exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    jobs = []
    for src in sources:
        obj = OBJ / (src.stem + ".o")
        if force or not _newer(obj, [src, *_deps(src), Path(__file__)]):
            cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append((src, cmd))

    def run(job):
        src, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, r in ex.map(run, jobs):
                if verbose or r.returncode != 0:
                    sys.stderr.write(f"--- {src.name}\n{r.stdout}{r.stderr}\n")
                if r.returncode != 0:
                    raise RuntimeError(f"nvcc failed on {src.name}")
    objs = [str(OBJ / (s.stem + ".o")) for s in sources]
    if force or jobs or not LIB.exists():
        tmp = LIB.with_suffix(".so.tmp")  # link beside the target, then rename: a reader never sees a half-written library
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(tmp), *objs,
               "-cudart", "static", "-Xlinker", "--no-undefined", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
        os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
