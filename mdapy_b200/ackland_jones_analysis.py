"""Ackland-Jones analysis, mirroring ``mdapy.ackland_jones_analysis.AcklandJonesAnalysis``
(src/mdapy/ackland_jones_analysis.py:13-107).  ``aja``: 0 other, 1 fcc, 2 hcp, 3 bcc, 4 ico."""
from __future__ import annotations

from typing import Optional

import numpy as np

from .box import Box
from .device import LIST_KNN, DeviceSystem
from .frame import Frame


class AcklandJonesAnalysis:
    def __init__(self, data, box: Box, verlet_list: Optional[np.ndarray] = None,
                 distance_list: Optional[np.ndarray] = None, dev: Optional[DeviceSystem] = None,
                 device: int = 0) -> None:
        self.data = Frame.from_any(data)
        self.box = box
        self.verlet_list = verlet_list
        self.distance_list = distance_list
        self._dev = dev
        self._device = device

    def compute(self) -> None:
        dev = self._dev
        if dev is None:
            dev = DeviceSystem(self._device)
            d, b = self.data, self.box
            dev.set_atoms(d["x"], d["y"], d["z"], b.box, b.origin, b.boundary)
            dev.put_neighbor(self.verlet_list, self.distance_list, kind=LIST_KNN)
        self.aja = dev.aja()
