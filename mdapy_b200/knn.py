"""k-nearest neighbours, mirroring ``mdapy.knn.NearestNeighbor`` (src/mdapy/knn.py:14-129)."""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import tool_function as tool
from .box import Box
from .device import DeviceSystem
from .frame import Frame

MAX_K = 24


class NearestNeighbor:
    def __init__(self, data, box: Box, k: int, device: int = 0):
        data = Frame.from_any(data)
        for col in ("x", "y", "z"):
            assert col in data.columns, f"data must contain column {col!r}."
        assert data.shape[0] > 0, "data must contain at least one atom."
        k = int(k)
        assert 1 <= k <= MAX_K, f"k must be in [1, {MAX_K}], got {k}."
        self.data = data
        self.box = box
        self.k = k
        self._device = device
        self.dev: Optional[DeviceSystem] = None
        self._host = None

    def _check_repeat_nearest(self):
        repeat = [1, 1, 1]
        N = self.data.shape[0]
        if self.k > N:
            assert sum(self.box.boundary) > 0, (
                f"Need periodic boundary if you want to query {self.k} neighbors "
                f"in {N}-atom system."
            )
            while np.prod(repeat) * N < self.k:
                for i in range(3):
                    if self.box.boundary[i] == 1:
                        repeat[i] += 3  # a safe number
        return repeat

    def compute(self, dev: Optional[DeviceSystem] = None, fetch: bool = True):
        data, box = self.data, self.box
        repeat = self._check_repeat_nearest()
        if sum(repeat) != 3:
            self._enlarge_data, self._enlarge_box = tool.replicate(data, box, *repeat)
            box, data = self._enlarge_box, self._enlarge_data
            dev = None
        if dev is None:
            dev = DeviceSystem(self._device)
            dev.set_atoms(data["x"], data["y"], data["z"], box.box, box.origin, box.boundary)
        self.dev = dev
        dev.build_knn(self.k)
        self._host = None
        if fetch:
            v, d, _ = dev.fetch_neighbor(True, True, False)
            self._host = (v, d)

    def _fetch(self):
        if self._host is None:
            if self.dev is None:
                raise AttributeError("call compute() first")
            v, d, _ = self.dev.fetch_neighbor(True, True, False)
            self._host = (v, d)
        return self._host

    @property
    def indices_py(self) -> np.ndarray:
        return self._fetch()[0]

    @property
    def distances_py(self) -> np.ndarray:
        return self._fetch()[1]
