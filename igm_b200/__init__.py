"""igm_b200 - the Hi-C A-step of IGM (and the assignment steps around it) on B200.

Public surface:

* ``igm_b200.ActdistEngine`` - ctypes front end of the C-ABI library ``libigmk.so``
  (include/igmk.h): coordinates staged once in HBM, then ``actdist`` / ``contact_counts`` /
  ``damid_actdist`` / ``restraint_select`` / ``sprite_rg2`` / ``rank_match``.
* ``igm_b200.steps`` - drop-ins for the reference's Step classes (same names and files).
* ``igm_b200.contact.get_simulated_hic`` - the call ``igm-report`` makes.

There is no CPU fallback: without the built library or a CUDA device every compute entry
point raises ``igm_b200.IgmkError``.
"""
from ._lib import IgmkError  # noqa: F401
from .engine import ActdistEngine  # noqa: F401

__version__ = "0.1.0"
__all__ = ["ActdistEngine", "IgmkError", "__version__"]
