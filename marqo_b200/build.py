"""In-tree build of libmarqo_b200.so (CUDA kernels + C ABI) and of the CPU oracle's C restatement.

nvcc cross-compiles sm_100a without a GPU.  The shared library has no torch dependency and links cudart
statically, so it loads in any process; it shares the primary CUDA context with torch when both are present.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libmarqo_b200.so"
ORACLE_DIR = REPO_ROOT / "oracle"
ORACLE_LIB = ORACLE_DIR / "libscore_oracle.so"

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O3",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libmarqo_b200.so")


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest(paths: list[Path]) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_native(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ (one object per file, parallel) and link libmarqo_b200.so."""
    if not force and os.environ.get("MARQO_B200_USE_PREBUILT") and LIB_PATH.exists():
        return LIB_PATH   # dev runs on the GPU box: use the library that travelled with the snapshot as it is
    srcs = _sources()
    deps = srcs + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.inc")) + [REPO_ROOT / "include" / "marqo_b200.h"]
    stamp = PKG_DIR / "build" / "stamp"
    digest = _digest(deps)
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB_PATH
    obj_dir = PKG_DIR / "build"
    obj_dir.mkdir(exist_ok=True)
    nvcc = _nvcc()
    procs = []
    for s in srcs:
        obj = obj_dir / (s.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(s), "-o", str(obj)]
        procs.append((s, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    log = []
    for s, obj, p in procs:
        out, _ = p.communicate()
        log.append(f"== {s.name}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s.name}:\n{out}")
        objs.append(str(obj))
    (obj_dir / "ptxas.log").write_text("\n".join(log))
    if verbose:
        print("\n".join(log))
    link = [nvcc, "-shared", "-o", str(LIB_PATH), *objs, "-gencode", "arch=compute_100a,code=sm_100a",
            "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    stamp.write_text(digest)
    return LIB_PATH


def build_oracle(force: bool = False) -> Path:
    """Compile the oracle's C restatement (test infrastructure only; never linked into the product)."""
    src = ORACLE_DIR / "score_oracle.c"
    if not force and ORACLE_LIB.exists() and ORACLE_LIB.stat().st_mtime >= src.stat().st_mtime:
        return ORACLE_LIB
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        raise RuntimeError("gcc not found; cannot build the C oracle")
    cmd = [cc, "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fopenmp", str(src), "-o",
           str(ORACLE_LIB), "-lm"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"oracle build failed:\n{r.stdout}")
    return ORACLE_LIB


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_native(force=force, verbose="-v" in sys.argv))
    print(build_oracle(force=force))
