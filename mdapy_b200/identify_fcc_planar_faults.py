"""FCC planar-fault identification, mirroring ``mdapy.identify_fcc_planar_faults.IdentifyFccPlanarFaults``
(src/mdapy/identify_fcc_planar_faults.py:58-84, kernel src/identify_fcc_planar_faults.cpp:43-241).
``fault_types``: 0 non-HCP, 1 other, 2 intrinsic stacking fault, 3 coherent twin boundary, 4 multi-layer
stacking fault, 5 extrinsic stacking fault.

``ptm_indices`` are the 12 matched neighbours per atom in the HCP template's point order; ``index_order`` says
whose: "mdapy_b200" (this library's PTM, the default) or "reference" (extern/ptm of mdapy)."""
from __future__ import annotations

import numpy as np

from . import _lib as L


class IdentifyFccPlanarFaults:
    def __init__(self, structure_types: np.ndarray, ptm_indices: np.ndarray, cal_esf: bool = True,
                 index_order: str = "mdapy_b200"):
        self.structure_types = L.i32(structure_types)
        self.ptm_indices = L.i32(np.ascontiguousarray(ptm_indices))
        assert self.ptm_indices.ndim == 2 and self.ptm_indices.shape[1] == 12
        assert index_order in ("mdapy_b200", "reference")
        self.cal_esf = bool(cal_esf)
        self.index_order = index_order

    def compute(self) -> None:
        n = self.structure_types.shape[0]
        self.fault_types = np.zeros(n, np.int32)
        L.check(L.lib().mdb_identify_sftb_fcc(None, 0, None, L.iptr(self.ptm_indices), L.iptr(self.structure_types), n,
                                              L.iptr(self.fault_types), int(self.cal_esf),
                                              0 if self.index_order == "mdapy_b200" else 1, 0))
