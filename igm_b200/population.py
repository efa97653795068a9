"""Population (.hss) and probability-matrix (.hcs) containers for the A-step.

Host-side replacements for the pieces of ``alabtools`` the hot path touches
(alabtools is a third-party dependency that is neither vendored in the
reference nor installed here; call sites:
igm/steps/ActivationDistanceStep.py:124-126,166,171,202,382-432).

* ``Population``  - what ``get_actdist`` reads from ``HssFile``: coordinates
  ``(nbead, nstruct, 3)`` float32 bead-major (igm/core/step.py:373,
  igm/_preprocess.py:102-105), ``radii``, ``index.chrom``,
  ``index.copy_index`` (haploid bin -> bead ids).
* ``ProbMatrix``  - what ``setup`` reads from ``Contactmatrix``: CSR of the
  strict upper triangle (``indptr, indices, data``) and ``index.chrom`` on
  haploid bins.
"""
from __future__ import annotations

import json
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import hdf5


class CopyIndex:
    """CSR form of ``index.copy_index``: haploid bin -> list of bead ids."""

    def __init__(self, ptr: np.ndarray, beads: np.ndarray):
        self.ptr = np.ascontiguousarray(ptr, dtype=np.int32)
        self.beads = np.ascontiguousarray(beads, dtype=np.int32)

    @classmethod
    def from_dict(cls, d: Dict) -> "CopyIndex":
        n = len(d)
        ptr = np.zeros(n + 1, dtype=np.int32)
        beads: List[int] = []
        for k in range(n):
            v = d[k] if k in d else d[str(k)]
            beads.extend(int(x) for x in v)
            ptr[k + 1] = len(beads)
        return cls(ptr, np.asarray(beads, dtype=np.int32))

    @classmethod
    def diploid(cls, n_hap: int, n_diploid: int) -> "CopyIndex":
        """First ``n_diploid`` bins have two copies ``[i, i + n_hap]`` (as in
        the demo file), the rest one copy ``[i]``."""
        ncopies = np.where(np.arange(n_hap) < n_diploid, 2, 1).astype(np.int32)
        ptr = np.concatenate([[0], np.cumsum(ncopies)]).astype(np.int32)
        beads = np.empty(ptr[-1], dtype=np.int32)
        beads[ptr[:-1]] = np.arange(n_hap)
        two = np.nonzero(ncopies == 2)[0]
        beads[ptr[two] + 1] = two + n_hap
        return cls(ptr, beads)

    def __len__(self):
        return len(self.ptr) - 1

    def __getitem__(self, i: int) -> List[int]:
        return [int(x) for x in self.beads[self.ptr[i]:self.ptr[i + 1]]]

    def ncopies(self) -> np.ndarray:
        return np.diff(self.ptr)

    def to_dict(self) -> Dict[int, List[int]]:
        return {i: self[i] for i in range(len(self))}


class Population:
    def __init__(self, coordinates: np.ndarray, radii: np.ndarray,
                 chrom: np.ndarray, copy_index: CopyIndex,
                 copy: Optional[np.ndarray] = None):
        coordinates = np.asarray(coordinates)
        if coordinates.ndim != 3 or coordinates.shape[2] != 3:
            raise ValueError("coordinates must be (nbead, nstruct, 3)")
        self.coordinates = np.ascontiguousarray(coordinates, dtype=np.float32)
        self.radii = np.ascontiguousarray(radii, dtype=np.float32)
        self.chrom = np.ascontiguousarray(chrom, dtype=np.int32)
        self.copy_index = copy_index
        self.copy = None if copy is None else np.ascontiguousarray(copy, dtype=np.int32)
        if self.radii.shape[0] != self.nbead or self.chrom.shape[0] != self.nbead:
            raise ValueError("radii/chrom length must equal nbead")

    @property
    def nbead(self) -> int:
        return self.coordinates.shape[0]

    @property
    def nstruct(self) -> int:
        return self.coordinates.shape[1]

    @property
    def n_hap(self) -> int:
        return len(self.copy_index)

    def chrom_hap(self) -> np.ndarray:
        """chrom[] as get_actdist indexes it: by *haploid* bin
        (ActivationDistanceStep.py:386,405) - valid because copy-0 beads come
        first in bead order."""
        return np.ascontiguousarray(self.chrom[:self.n_hap])

    @classmethod
    def from_hss(cls, path: str) -> "Population":
        with hdf5.open_h5(path) as f:
            crd = np.asarray(f["coordinates"][:], dtype=np.float32)
            radii = np.asarray(f["radii"][:], dtype=np.float32)
            chrom = np.asarray(f["index"]["chrom"][:], dtype=np.int32)
            ci = f["index"]["copy_index"][()]
            if isinstance(ci, np.ndarray):            # fixed-length string dataset (save_hss)
                ci = ci.tobytes() if ci.dtype.kind in "SV" else ci.item()
            if isinstance(ci, (bytes, np.bytes_)):
                ci = bytes(ci).rstrip(b"\x00").decode("utf-8")
            copy = np.asarray(f["index"]["copy"][:], dtype=np.int32)
        return cls(crd, radii, chrom, CopyIndex.from_dict(json.loads(ci)), copy)

    def save_hss(self, path: str) -> None:
        """Writes the subset of the .hss layout this package reads back
        (igm/_preprocess.py:90-110).  copy_index is stored as a fixed-length
        JSON string (the reference stores a variable-length string)."""
        ci = json.dumps({str(k): v for k, v in self.copy_index.to_dict().items()})
        copy = self.copy if self.copy is not None else np.zeros(self.nbead, np.int32)
        hdf5.write_h5(path, {
            "coordinates": self.coordinates,
            "radii": self.radii,
            "index/chrom": self.chrom,
            "index/copy": copy,
            "index/copy_index": np.array(ci.encode("utf-8")),
        }, attrs={"version": np.int32(2), "nbead": np.int64(self.nbead),
                  "nstruct": np.int64(self.nstruct)})


class ProbMatrix:
    """Input Hi-C probability matrix (.hcs): strict-upper-triangle CSR."""

    def __init__(self, indptr, indices, data, chrom, diagonal=None):
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32)
        self.data = np.ascontiguousarray(data, dtype=np.float32)
        self.chrom = np.ascontiguousarray(chrom, dtype=np.int32)
        self.diagonal = diagonal
        self.n = len(self.indptr) - 1
        self.shape = (self.n, self.n)

    @classmethod
    def from_hcs(cls, path: str) -> "ProbMatrix":
        with hdf5.open_h5(path) as f:
            m = f["matrix"]
            indptr = np.asarray(m["indptr"][:])
            indices = np.asarray(m["indices"][:])
            data = np.asarray(m["data"][:])
            diag = np.asarray(m["diagonal"][:]) if "diagonal" in m else None
            chrom = np.asarray(f["index"]["chrom"][:])
        return cls(indptr, indices, data, chrom, diag)

    def save_hcs(self, path: str) -> None:
        diag = self.diagonal if self.diagonal is not None else np.ones(self.n, np.float32)
        hdf5.write_h5(path, {
            "matrix/indptr": self.indptr.astype(np.int32),
            "matrix/indices": self.indices,
            "matrix/data": self.data,
            "matrix/diagonal": np.asarray(diag, dtype=np.float32),
            "index/chrom": self.chrom,
        }, attrs={"nbin": np.int64(self.n), "resolution": np.int64(-1)})

    def rows(self) -> np.ndarray:
        """Row index of every stored non-zero, CSR (row-major) order - the
        order ``coo_generator`` yields them in (ActivationDistanceStep.py:171)."""
        return np.repeat(np.arange(self.n, dtype=np.int32), np.diff(self.indptr))
