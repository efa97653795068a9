"""In-tree build of libigmk.so for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libigmk.so")
SOURCES = ["igmk.cu"]
HEADERS = ["igmk_device.cuh", "igmk_actdist.cuh", "igmk_actdist_list.cuh", "igmk_actdist_slab.cuh", "igmk_contact.cuh", "igmk_restraint.cuh", "igmk_sprite.cuh",
           "igmk_rank.cuh", "../../include/igmk.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # bit-exact float32/float64 (no FMA contraction)
    "-shared", "-Xcompiler", "-fPIC",
    "-cudart", "static",
]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: str = OUT, defines=()) -> str:
    """``out`` / ``defines`` build a tuning variant next to the shipped library
    (loaded through IGMK_LIB_PATH); the default call builds igm_b200/libigmk.so."""
    if out == OUT and not defines and not force and not _stale():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + list(defines) + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libigmk.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return out


if __name__ == "__main__":
    _out = OUT
    if "--out" in sys.argv:
        _out = os.path.abspath(sys.argv[sys.argv.index("--out") + 1])
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=_out,
                defines=[a for a in sys.argv[1:] if a.startswith("-D")]))
