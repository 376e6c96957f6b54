"""Diamond structure identification, mirroring
``mdapy.identify_diamond_structure.IdentifyDiamondStructure`` (src/mdapy/identify_diamond_structure.py:15-124,
kernel src/cna.cpp:163-287).  ``pattern``: 0 other, 1 cubic diamond, 2 / 3 its 1st / 2nd neighbours,
4 hexagonal diamond, 5 / 6 its 1st / 2nd neighbours."""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import tool_function as tool
from .box import Box
from .device import LIST_KNN, DeviceSystem
from .frame import Frame
from .knn import NearestNeighbor


class IdentifyDiamondStructure:
    def __init__(self, data, box: Box, verlet_list: Optional[np.ndarray] = None,
                 dev: Optional[DeviceSystem] = None, device: int = 0):
        self.data = Frame.from_any(data)
        self.box = box
        self.verlet_list = verlet_list
        self.pattern = np.array([], dtype=np.int32)
        self._dev = dev          # device system that already holds atoms + a sorted list (>= 4 per row)
        self._device = device

    def compute(self):
        N = self.data.shape[0]
        if sum(self.box.boundary) == 0 and N <= 4:
            self.pattern = np.zeros(N, dtype=np.int32)
            return
        box, data = self.box, self.data
        dev = self._dev
        safe_L = 15  # identify_diamond_structure.py:94

        if dev is None and self.verlet_list is None:
            repeat = np.ceil(safe_L / self.box.get_thickness()).astype(int)
            for i in range(3):
                if self.box.boundary[i] == 0:
                    repeat[i] = 1
            if sum(repeat) != 3:
                data, box = tool._replicate_pos(data, box, *repeat)
            knn = NearestNeighbor(data, box, 4, device=self._device)
            knn.compute(fetch=False)
            dev = knn.dev
        elif dev is None:
            dev = DeviceSystem(self._device)
            dev.set_atoms(data["x"], data["y"], data["z"], box.box, box.origin, box.boundary)
            dev.put_neighbor(self.verlet_list, None, None, rc=-1.0, kind=LIST_KNN)
        self.pattern = dev.ids()
