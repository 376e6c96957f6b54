"""Common neighbour analysis, mirroring ``mdapy.common_neighbor_analysis.CommonNeighborAnalysis``
(src/mdapy/common_neighbor_analysis.py:17-154): fixed cut-off (``rc`` given -> _cna.fcna) or adaptive
(``rc=None`` -> 14-NN + _cna.acna).  ``pattern``: 0 other, 1 fcc, 2 hcp, 3 bcc, 4 ico."""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import tool_function as tool
from .box import Box
from .device import LIST_CUTOFF, LIST_KNN, DeviceSystem
from .frame import Frame
from .knn import NearestNeighbor
from .neighbor import Neighbor


class CommonNeighborAnalysis:
    def __init__(self, data, box: Box, verlet_list: Optional[np.ndarray] = None,
                 neighbor_number: Optional[np.ndarray] = None, rc: Optional[float] = None,
                 dev: Optional[DeviceSystem] = None, device: int = 0):
        self.data = Frame.from_any(data)
        self.box = box
        self.verlet_list = verlet_list
        self.neighbor_number = neighbor_number
        if rc is not None:
            assert rc > 0
        self.rc = rc
        self.pattern = None
        self._dev = dev          # device system that already holds atoms + the list to use
        self._device = device

    def compute(self):
        N = self.data.shape[0]
        if sum(self.box.boundary) == 0 and N <= 14:
            self.pattern = np.zeros(N, dtype=np.int32)
            return
        box, data = self.box, self.data
        dev = self._dev
        wrap_pos_L = 15  # common_neighbor_analysis.py:101

        if dev is None and self.verlet_list is None:
            repeat = np.ceil(wrap_pos_L / self.box.get_thickness()).astype(int)
            for i in range(3):
                if self.box.boundary[i] == 0:
                    repeat[i] = 1
            if sum(repeat) != 3:
                data, box = tool._replicate_pos(data, box, *repeat)
            if self.rc is None:
                knn = NearestNeighbor(data, box, 14, device=self._device)
                knn.compute(fetch=False)
                dev = knn.dev
            else:
                repeat = box.check_small_box(self.rc)
                if sum(repeat) != 3:
                    data, box = tool._replicate_pos(data, box, *repeat)
                neigh = Neighbor(self.rc, box, data, device=self._device)
                neigh.compute(fetch=False)
                dev = neigh.dev
        elif dev is None:
            assert self.rc is None or self.neighbor_number is not None
            dev = DeviceSystem(self._device)
            dev.set_atoms(data["x"], data["y"], data["z"], box.box, box.origin, box.boundary)
            # a caller-supplied list carries no distance bound the fast bond test could rely on: rc=-1
            dev.put_neighbor(self.verlet_list, None, self.neighbor_number, rc=-1.0,
                             kind=LIST_CUTOFF if self.rc is not None else LIST_KNN)
        self.pattern = dev.acna() if self.rc is None else dev.fcna(self.rc)
