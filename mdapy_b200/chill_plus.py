"""CHILL+ water-phase identification, mirroring ``mdapy.chill_plus.ChillPlus`` (src/mdapy/chill_plus.py:58-116,
kernel src/chill_plus.cpp:76-181).  ``pattern``: 0 other, 1 hexagonal ice, 2 cubic ice, 3 interfacial ice,
4 gas hydrate, 5 interfacial gas hydrate.  The system must hold the molecule centres only (oxygens)."""
from __future__ import annotations

from typing import Optional

import numpy as np

from .box import Box
from .device import LIST_CUTOFF, DeviceSystem
from .frame import Frame
from .neighbor import Neighbor


class ChillPlus:
    def __init__(self, data, box: Box, cutoff: float = 3.5, verlet_list: Optional[np.ndarray] = None,
                 distance_list: Optional[np.ndarray] = None, neighbor_number: Optional[np.ndarray] = None,
                 dev: Optional[DeviceSystem] = None, device: int = 0):
        self.data = Frame.from_any(data)
        self.box = box
        self.cutoff = float(cutoff)
        self.verlet_list, self.distance_list, self.neighbor_number = verlet_list, distance_list, neighbor_number
        self.pattern = np.array([], dtype=np.int32)
        self._dev, self._device = dev, device

    def compute(self) -> None:
        dev = self._dev
        if dev is None:
            if self.verlet_list is None or self.distance_list is None or self.neighbor_number is None:
                neigh = Neighbor(self.cutoff, self.box, self.data, device=self._device)
                neigh.compute(fetch=False)
                dev = neigh.dev
            else:
                dev = DeviceSystem(self._device)
                dev.set_atoms(self.data["x"], self.data["y"], self.data["z"], self.box.box, self.box.origin,
                              self.box.boundary)
                dev.put_neighbor(self.verlet_list, self.distance_list, self.neighbor_number, rc=self.cutoff,
                                 kind=LIST_CUTOFF)
        self.pattern = dev.chill_plus(self.cutoff)
