"""Local structural (pair) entropy, mirroring ``mdapy.structure_entropy.StructureEntropy``
(src/mdapy/structure_entropy.py:14-145; kernel: src/structure_entropy.cpp:11-103)."""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _lib as L
from .box import Box
from .device import DeviceSystem


class StructureEntropy:
    def __init__(self, box: Box, verlet_list: Optional[np.ndarray] = None, distance_list: Optional[np.ndarray] = None,
                 neighbor_number: Optional[np.ndarray] = None, rc: float = 5.0, sigma: float = 0.2,
                 use_local_density: bool = False, average_rc: float = 0.0, dev: Optional[DeviceSystem] = None) -> None:
        self.box = box
        self.verlet_list = verlet_list
        self.distance_list = distance_list
        self.neighbor_number = neighbor_number
        self.rc = rc
        self.sigma = sigma
        self.use_local_density = use_local_density
        self.average_rc = average_rc
        self._dev = dev

    def compute(self):
        if self.average_rc > 0:
            assert self.average_rc <= self.rc, "average_rc should be smaller than rc."
        if self._dev is not None:
            ent, ave = self._dev.structure_entropy(self.rc, self.sigma, self.use_local_density, self.box.volume,
                                                   self.average_rc)
            self.entropy = ent
            if self.average_rc > 0:
                self.entropy_ave = ave
            return
        d, n = L.f64(self.distance_list), L.i32(self.neighbor_number)
        N, M = d.shape
        self.entropy = np.zeros(N)
        L.check(L.lib().mdb_calculate_structure_entropy(float(self.rc), float(self.sigma), int(bool(self.use_local_density)),
                                                        float(self.box.volume), L.dptr(d), N, M, L.iptr(n),
                                                        L.dptr(self.entropy), 1))
        if self.average_rc > 0:
            v = L.i32(self.verlet_list)
            self.entropy_ave = np.zeros_like(self.entropy)
            L.check(L.lib().mdb_average_by_neighbor(float(self.average_rc), L.iptr(v), N, M, L.dptr(d), L.iptr(n),
                                                    L.dptr(self.entropy), L.dptr(self.entropy_ave), 1, 1))
