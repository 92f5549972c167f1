"""Builds libvdfgpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIBDIR = ROOT / "lib"
LIB = LIBDIR / "libvdfgpu.so"
SOURCES = ["api_core.cu", "api_r1cs.cu", "api_sumcheck.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _newer(target: Path, deps) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(d).stat().st_mtime <= t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags=(), tag: str = "") -> Path:
    """tag != "" builds an experiment variant lib/libvdfgpu_<tag>.so with extra nvcc flags (tools/ only)."""
    global LIB
    LIBDIR.mkdir(exist_ok=True)
    lib = LIBDIR / (f"libvdfgpu_{tag}.so" if tag else "libvdfgpu.so")
    headers = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.hpp")) + [ROOT.parent / "include" / "vdfgpu.h"]
    objs = []
    procs = []
    for src in SOURCES:
        obj = LIBDIR / (src[:-3] + (f"_{tag}" if tag else "") + ".o")
        objs.append(obj)
        if not force and _newer(obj, [CSRC / src] + headers):
            continue
        cmd = [_nvcc(), *NVCC_FLAGS, *extra_flags, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    if force or procs or not _newer(lib, objs):
        cmd = [_nvcc(), "-shared", "-o", str(lib), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}")
    return lib


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--tag", default="")
    ap.add_argument("--flag", action="append", default=[])
    a = ap.parse_args()
    print(build(force=a.force, verbose=True, extra_flags=a.flag, tag=a.tag))
