"""Build libagpknn.so (hand-written sm_100a CUDA kernels + C ABI) in-tree with nvcc.

``python -m agplace_b200.build`` compiles every translation unit under ``csrc/`` for
``-gencode arch=compute_100a,code=sm_100a`` (objects in ``csrc/_obj``, in parallel) and links
``agplace_b200/libagpknn.so``.  nvcc cross-compiles without a GPU, so this also runs on the
CPU-only build box; the resulting ``.so`` travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = CSRC / "_obj"
LIB = PKG / "libagpknn.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]
if os.environ.get("AGP_BUILD_DEBUG", "") not in ("", "0"):
    # development build: AGP_SCREEN_* / AGP_TC_* environment seeding of the per-index knobs and the result-changing
    # bandwidth probes (skip_epi, skip_mma).  The product build has none of it.
    CFLAGS.append("-DAGP_DEBUG_KNOBS")

# (source, define value or None, object name)
UNITS = (
    [("agpknn.cu", None, "agpknn.o"), ("k_misc.cu", None, "k_misc.o")]
    + [("k_tc.cu", e, f"k_tc_{e}.o") for e in (2, 4, 8, 16)]
    + [("k_screen.cu", e, f"k_screen_{e}.o") for e in (8, 16, 32)]
    + [("k_select.cu", e, f"k_select_{e}.o") for e in (2, 4, 8, 16, 32)]
)


def _deps_hash(src: Path, define) -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [src, PKG.parent / "include" / "agpknn.h"]):
        h.update(f.read_bytes())
    h.update(repr((define, ARCH, CFLAGS)).encode())
    return h.hexdigest()


def _compile(unit, verbose=False):
    src, define, obj = unit
    out = OBJ / obj
    stamp = OBJ / (obj + ".hash")
    want = _deps_hash(CSRC / src, define)
    if out.exists() and stamp.exists() and stamp.read_text() == want:
        return obj, "cached", ""
    cmd = [NVCC, *ARCH, *CFLAGS, "-c", str(CSRC / src), "-o", str(out)]
    if define is not None:
        cmd.insert(1, f"-DAGP_E={define}")
    if verbose:
        cmd += ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src} (AGP_E={define}):\n{r.stdout}\n{r.stderr}")
    stamp.write_text(want)
    return obj, "built", r.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    if force:
        for f in OBJ.glob("*.hash"):
            f.unlink()
    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 4)) as ex:
        results = list(ex.map(lambda u: _compile(u, verbose), UNITS))
    rebuilt = [r for r in results if r[1] == "built"]
    if verbose:
        for obj, _, log in rebuilt:
            print(f"== {obj}\n{log}")
    if rebuilt or not LIB.exists():
        objs = [str(OBJ / u[2]) for u in UNITS]
        cmd = [NVCC, *ARCH, "-shared", "-o", str(LIB), *objs, "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(lib)
