"""In-tree build of the CUDA library (``libdajin2_b200.so``) for sm_100a.

``nvcc`` cross-compiles without a GPU, so this runs in the build container; the resulting ``.so`` is git-ignored
but travels to the GPU box with the repository snapshot.
"""

from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libdajin2_b200.so"
SOURCES = ["cabi.cu", "encode.cu", "encode_v1.cu", "hist.cu", "kmer.cu", "cluster.cu", "readstat.cu", "cstag.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found: the dajin2_b200 CUDA library cannot be built")
    return exe


def needs_build() -> bool:
    if not LIB.exists():
        return True
    stamp = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + sorted(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "dajin2_b200.h"]
    return any(d.stat().st_mtime > stamp for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    objdir = HERE / "build"
    objdir.mkdir(exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = objdir / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    tmp = LIB.with_suffix(".so.tmp%d" % os.getpid())
    subprocess.run([nvcc, "-shared", "--cudart", "shared", "-Wno-deprecated-gpu-targets", "-o", str(tmp), *objs], check=True)
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
