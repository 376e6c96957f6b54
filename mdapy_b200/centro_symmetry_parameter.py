"""Centro-symmetry parameter, mirroring ``mdapy.centro_symmetry_parameter.CentroSymmetryParameter``
(src/mdapy/centro_symmetry_parameter.py:13-107)."""
from __future__ import annotations

from typing import Optional

import numpy as np

from .box import Box
from .device import LIST_KNN, DeviceSystem
from .frame import Frame


class CentroSymmetryParameter:
    def __init__(self, data, box: Box, N: int, verlet_list: Optional[np.ndarray] = None,
                 dev: Optional[DeviceSystem] = None, device: int = 0) -> None:
        self.data = Frame.from_any(data)
        self.box = box
        assert N % 2 == 0 and N > 0, f"N must be a positive even number: {N}."
        self.N = int(N)
        self.verlet_list = verlet_list
        self._dev = dev
        self._device = device

    def compute(self) -> None:
        dev = self._dev
        if dev is None:
            dev = DeviceSystem(self._device)
            d, b = self.data, self.box
            dev.set_atoms(d["x"], d["y"], d["z"], b.box, b.origin, b.boundary)
            dev.put_neighbor(self.verlet_list, kind=LIST_KNN)
        self.csp = dev.csp(self.N)
