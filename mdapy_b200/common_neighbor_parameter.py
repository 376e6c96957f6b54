"""Common neighbour parameter, mirroring ``mdapy.common_neighbor_parameter.CommonNeighborParameter``
(src/mdapy/common_neighbor_parameter.py:11-117; kernel: src/common_neighbor_parameter.cpp:10-136)."""
from __future__ import annotations

from typing import Optional

import numpy as np

from .box import Box
from .device import LIST_CUTOFF, DeviceSystem
from .frame import Frame


class CommonNeighborParameter:
    def __init__(self, data, box: Box, rc: float, verlet_list: Optional[np.ndarray] = None,
                 distance_list: Optional[np.ndarray] = None, neighbor_number: Optional[np.ndarray] = None,
                 dev: Optional[DeviceSystem] = None, device: int = 0) -> None:
        self.data = Frame.from_any(data)
        self.box = box
        self.rc = rc
        assert rc > 0
        self.verlet_list = verlet_list
        self.distance_list = distance_list
        self.neighbor_number = neighbor_number
        self._dev = dev
        self._device = device

    def compute(self) -> None:
        dev = self._dev
        if dev is None:
            dev = DeviceSystem(self._device)
            d, b = self.data, self.box
            dev.set_atoms(d["x"], d["y"], d["z"], b.box, b.origin, b.boundary)
            dev.put_neighbor(self.verlet_list, self.distance_list, self.neighbor_number, rc=float(self.rc),
                             kind=LIST_CUTOFF)
        self.cnp = dev.cnp(float(self.rc))
