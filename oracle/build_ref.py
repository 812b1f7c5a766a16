"""Recipe for ``oracle/_ref``: the UNMODIFIED reference package, importable on the GPU box.

TEST INFRASTRUCTURE — see the header of ``oracle/dajin2_oracle.py``.  The reference (akikuno/DAJIN2 v0.9.3) is pure
Python, so "building" it is a file copy: ``python oracle/build_ref.py`` copies ``/root/reference/src/DAJIN2/**/*.py``
to ``oracle/_ref/DAJIN2`` byte for byte and writes ``oracle/_ref/MANIFEST.json`` (SHA-256 of every file).  The output
directory is git-ignored (no reference source ever enters the history) but NOT gpurun-ignored, so it travels to the
GPU box like the built ``.so`` files; ``/root/reference`` itself does not exist there.  The third-party modules the
reference imports but this image lacks (``mappy``, ``midsv``, ``pysam``, ``cstag``, ``rapidfuzz``, ``wslPath``) are the
empty stand-ins under ``oracle/ref_stubs`` (ours, tracked); none of them is touched by the hot path.

``load()`` puts both on ``sys.path`` and returns the imported package; only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU arms may call it.
"""

from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
SOURCE = Path("/root/reference/src/DAJIN2")
TARGET = HERE / "_ref"
STUBS = HERE / "ref_stubs"


def available() -> bool:
    return (TARGET / "DAJIN2" / "__init__.py").exists() and (TARGET / "MANIFEST.json").exists()


def build(force: bool = False) -> Path | None:
    """Copy the reference's Python sources (no templates / static assets of the GUI) when the checkout is present."""
    if not SOURCE.exists():
        return TARGET if available() else None
    files = sorted(p for p in SOURCE.rglob("*.py"))
    manifest = {str(p.relative_to(SOURCE)): hashlib.sha256(p.read_bytes()).hexdigest() for p in files}
    stamp = TARGET / "MANIFEST.json"
    if not force and stamp.exists():
        try:
            if json.loads(stamp.read_text()).get("files") == manifest:
                return TARGET
        except ValueError:
            pass
    tmp = TARGET.with_name("_ref.tmp%d" % os.getpid())
    shutil.rmtree(tmp, ignore_errors=True)
    for p in files:
        dst = tmp / "DAJIN2" / p.relative_to(SOURCE)
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(p, dst)
    version = "unknown"
    pyproject = SOURCE.parent.parent / "pyproject.toml"
    if pyproject.exists():
        for line in pyproject.read_text().splitlines():
            if line.startswith("version"):
                version = line.split("=", 1)[1].strip().strip('"')
                break
    (tmp / "MANIFEST.json").write_text(json.dumps(
        {"package": "DAJIN2", "version": version, "source": str(SOURCE), "files": manifest}, indent=1, sort_keys=True))
    shutil.rmtree(TARGET, ignore_errors=True)
    os.replace(tmp, TARGET)
    return TARGET


def load():
    """Import the reference (``oracle/_ref`` first, the read-only checkout as a fall-back in the build container)
    with BLAS / OpenMP pinned to one thread as its ``main.py:5-9`` does.  Returns the ``DAJIN2`` package."""
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.setdefault(v, "1")
    if "DAJIN2" in sys.modules:
        return sys.modules["DAJIN2"]
    if available():
        root = str(TARGET)
    elif SOURCE.exists():
        root = str(SOURCE.parent)
    else:
        raise ImportError("the reference is not available: run `python oracle/build_ref.py` where /root/reference exists")
    for p in (root, str(STUBS)):
        if p not in sys.path:
            sys.path.insert(0, p)
    import DAJIN2  # noqa: F401

    return sys.modules["DAJIN2"]


def version() -> str:
    try:
        return json.loads((TARGET / "MANIFEST.json").read_text())["version"]
    except (OSError, ValueError, KeyError):
        return "unknown"


if __name__ == "__main__":
    out = build(force="--force" in sys.argv)
    print("oracle/_ref:", out, "available" if available() else "NOT available")
