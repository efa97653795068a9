"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports
every symbol include/igmk.h declares; without a GPU every compute entry point
fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests import helpers as H


def _header_symbols():
    txt = open(os.path.join(H.ROOT, "include", "igmk.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(igmk_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from igm_b200 import _lib, build
    build.build()
    lib = _lib.load()
    syms = _header_symbols()
    assert set(syms) == set(_lib.EXPORTS)
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.igmk_version() == 100


def test_result_struct_layout():
    from igm_b200 import _lib
    dt = _lib.PAIR_RESULT_DTYPE
    assert dt.itemsize == 32
    assert [dt.fields[n][1] for n in ("d2_sel_bits", "contact_count", "o", "nrec", "p", "dist", "prob")] == \
        [0, 4, 8, 12, 16, 24, 28]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from igm_b200 import _lib
    from igm_b200.engine import ActdistEngine
    with pytest.raises(_lib.IgmkError) as e:
        ActdistEngine(nbead=4, nstruct=4, device=0)
    assert e.value.code == _lib.IGMK_ECUDA


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (parity would be void)."""
    pkg = os.path.join(H.ROOT, "igm_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "oracle/" not in src and "oracle." not in src, f
