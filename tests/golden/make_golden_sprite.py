#!/usr/bin/env python
"""Golden vectors for the SPRITE assignment arithmetic, produced by the reference's OWN
compute_gyration_radius (igm/cython_compiled/sprite.pyx:104-283, which drives its native
get_rg2s_cpp): sprite.pyx and cpp_sprite_assignment.cpp are compiled unmodified from
/root/reference in a scratch directory (Cython + g++, the reference's own setup.py recipe)
and run on the populations of tests/golden/damid_small.npz.  The representative segment of
each chromosome is drawn with np.random.choice (:222-223), so every cluster is evaluated
after np.random.seed(seed0 + k) and the seed is stored with the result.
Runs only in the build container (/root/reference present):
    python tests/golden/make_golden_sprite.py
"""
import importlib
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("IGM_REFERENCE_ROOT", "/root/reference")
SEED0 = 4100


def build_reference_module():
    tmp = tempfile.mkdtemp(prefix="igm_sprite_ref_")
    src = os.path.join(REF, "igm", "cython_compiled")
    for f in ("sprite.pyx", "cpp_sprite_assignment.cpp", "cpp_sprite_assignment.h"):
        shutil.copy(os.path.join(src, f), tmp)
    with open(os.path.join(tmp, "setup.py"), "w") as fh:
        fh.write("from setuptools import setup, Extension\nfrom Cython.Build import cythonize\nimport numpy\n"
                 "setup(name='sprite', ext_modules=cythonize([Extension('sprite', ['sprite.pyx', "
                 "'cpp_sprite_assignment.cpp'], language='c++', include_dirs=[numpy.get_include()])], "
                 "language_level=3))\n")
    subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=tmp, check=True,
                   capture_output=True)
    sys.path.insert(0, tmp)
    return importlib.import_module("sprite")


class _Index:
    def __init__(self, chrom, copy_index):
        self.chrom = chrom
        self.copy_index = copy_index


def make_clusters(rng, chrom_hap, n):
    """Segment sets of 2..12 loci: single-chromosome, few-chromosome and genome-wide ones."""
    out = []
    chroms = np.unique(chrom_hap)
    for k in range(n):
        kind = k % 3
        if kind == 0:                                  # one chromosome
            c = rng.choice(chroms)
            pool = np.nonzero(chrom_hap == c)[0]
        elif kind == 1:                                # two or three chromosomes
            cs = rng.choice(chroms, size=min(len(chroms), int(rng.integers(2, 4))), replace=False)
            pool = np.nonzero(np.isin(chrom_hap, cs))[0]
        else:
            pool = np.arange(len(chrom_hap))
        size = int(min(len(pool), rng.integers(2, 13)))
        out.append(np.sort(rng.choice(pool, size=size, replace=False)).astype(np.int32))
    return out


def main():
    ref = build_reference_module()
    g = np.load(os.path.join(HERE, "damid_small.npz"))
    out = {"names": np.array(["n37", "n100", "n257"]), "seed0": np.int64(SEED0)}
    for name in ("n37", "n100", "n257"):
        crd = np.ascontiguousarray(g[name + "_coords"], np.float32)
        ptr, beads = g[name + "_copy_ptr"], g[name + "_copy_beads"]
        copy_index = {i: [int(b) for b in beads[ptr[i]:ptr[i + 1]]] for i in range(len(ptr) - 1)}
        chrom = np.asarray(g[name + "_chrom"])
        chrom_hap = chrom[[copy_index[i][0] for i in range(len(ptr) - 1)]]
        rng = np.random.default_rng(len(name) + crd.shape[1])
        clusters = make_clusters(rng, chrom_hap, 18)
        out[name + "_cluster_ptr"] = np.concatenate([[0], np.cumsum([len(c) for c in clusters])]).astype(np.int32)
        out[name + "_cluster_data"] = np.concatenate(clusters).astype(np.int32)
        rg_all, best_all, sel_all = [], [], []
        for k, cl in enumerate(clusters):
            np.random.seed(SEED0 + k)
            rg2s, best, sel = ref.compute_gyration_radius(crd, cl, _Index(chrom, copy_index), copy_index)
            rg_all.append(np.asarray(rg2s, np.float32))
            best_all.append(int(best))
            sel_all.append(np.asarray(sel, np.int32).reshape(-1))
        out[name + "_rg2s"] = np.stack(rg_all)
        out[name + "_best"] = np.asarray(best_all, np.int32)
        out[name + "_selected"] = np.concatenate(sel_all)          # cluster k: (nstruct, len(cluster k)) row-major
    np.savez_compressed(os.path.join(HERE, "sprite_small.npz"), **out)
    print("wrote sprite_small.npz:", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
