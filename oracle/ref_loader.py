"""TEST INFRASTRUCTURE ONLY - loader for the *unmodified* Python reference.

Imports the reference's own ``get_actdist`` functions from ``/root/reference``
(read-only, present only in the build container, never on the GPU box) so that
``tests/golden/make_golden.py`` can generate golden vectors and so that the
restatement in ``oracle/actdist_oracle.py`` can be pinned against it.

The reference package imports third-party modules that are not installed here
(``alabtools``, ``h5py``, ``ipyparallel``, ``zmq``, ``cloudpickle``, ``tornado``,
``matplotlib``, ...; igm/__init__.py:5-13, igm/steps/__init__.py:2-11,
igm/core/step.py:14-16).  ``get_actdist`` itself needs none of them - it only
duck-types its ``hss`` argument (igm/steps/ActivationDistanceStep.py:382-432) -
so a ``sys.meta_path`` finder hands out empty stub modules for the missing
imports and the arithmetic runs unmodified.

Nothing in the product path (``igm_b200``) may import this module.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
# the read-only reference tree in the build container; on the GPU box the copy of its
# Python package that `make -C oracle ref_py` left in git-ignored oracle/_ref/
REFERENCE_ROOT = os.environ.get("IGM_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REFERENCE_ROOT, "igm", "steps")):
    REFERENCE_ROOT = os.path.join(HERE, "_ref")

_STUB_ROOTS = (
    "alabtools", "h5py", "ipyparallel", "zmq", "cloudpickle", "tornado",
    "matplotlib", "dask", "distributed", "lammps", "tqdm_missing",
)


class _StubModule(types.ModuleType):
    """Module whose every attribute is another stub (callable, subclassable)."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        val = type(name, (object,), {"__init__": lambda self, *a, **k: None})
        setattr(self, name, val)
        return val


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        root = fullname.split(".")[0]
        if root in _STUB_ROOTS or fullname == "igm.cython_compiled.sprite":
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = False


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "igm", "steps"))


def install() -> None:
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    sys.meta_path.insert(0, _StubFinder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def load_reference():
    """Returns a dict of the reference's own callables for the hot path."""
    install()
    lb = importlib.import_module("igm.steps.ActivationDistanceStep")
    gp = importlib.import_module("igm.steps.GP_activation")
    ut = importlib.import_module("igm.utils.actdist")
    return {
        "lb_get_actdist": lb.get_actdist,          # ActivationDistanceStep.py:336
        "gp_get_actdist": gp.get_actdist,          # GP_activation.py:317
        "utils_get_actdist": ut.get_actdist,       # utils/actdist.py:15
        "cleanProbability": lb.cleanProbability,   # ActivationDistanceStep.py:314
        "actdist_fmt_str": lb.actdist_fmt_str,     # ActivationDistanceStep.py:38
        "actdist_shape": lb.actdist_shape,         # ActivationDistanceStep.py:32
        "ActivationDistanceStep": lb.ActivationDistanceStep,
    }


class FakeIndex:
    def __init__(self, chrom, copy_index):
        self.chrom = chrom
        self.copy_index = copy_index


class FakeHss:
    """Duck-typed stand-in for alabtools.analysis.HssFile: the five accessors
    ``get_actdist`` uses (ActivationDistanceStep.py:382-432)."""

    def __init__(self, coordinates, radii, chrom, copy_index):
        # coordinates: (nbead, nstruct, 3) float32, bead-major as in the .hss
        self.coordinates = coordinates
        self.radii = radii
        self.index = FakeIndex(chrom, copy_index)

    def get_nstruct(self):
        return self.coordinates.shape[1]

    def get_index(self):
        return self.index

    def get_radii(self):
        return self.radii

    def get_bead_crd(self, k):
        return self.coordinates[k]

    def close(self):
        pass
