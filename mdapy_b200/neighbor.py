"""Cut-off neighbour list, mirroring ``mdapy.neighbor.Neighbor`` (src/mdapy/neighbor.py:14-142)."""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import tool_function as tool
from .box import Box
from .device import DeviceSystem
from .frame import Frame


class Neighbor:
    """Same constructor, attributes and errors as the reference class.  After ``compute()``:
    ``verlet_list`` int32[N,M] (-1 padded), ``distance_list`` f64[N,M] (rc+1 padded),
    ``neighbor_number`` int32[N]; ``_enlarge_data/_enlarge_box`` when the box was replicated.
    The list also stays on the device in ``self.dev`` for chained descriptors."""

    def __init__(self, rc: float, box: Box, data, max_neigh: Optional[int] = None, device: int = 0):
        rc = float(rc)
        assert rc > 0, f"rc must be positive, got {rc}."
        if max_neigh is not None:
            max_neigh = int(max_neigh)
            assert max_neigh > 0, f"max_neigh must be positive, got {max_neigh}."
        data = Frame.from_any(data)
        for col in ("x", "y", "z"):
            assert col in data.columns, f"data must contain column {col!r}."
        self.rc = rc
        self.box = box
        self.data = data
        self.max_neigh = max_neigh
        self.N = self.data.shape[0]
        assert self.N > 0, "data must contain at least one atom."
        self._device = device
        self.dev: Optional[DeviceSystem] = None
        self._host = [None, None, None]

    def compute(self, dev: Optional[DeviceSystem] = None, fetch: bool = True):
        repeat = self.box.check_small_box(self.rc)
        if sum(repeat) != 3:
            self._enlarge_data, self._enlarge_box = tool.replicate(self.data, self.box, *repeat)
            data, box = self._enlarge_data, self._enlarge_box
            dev = None  # atoms on the device no longer match
        else:
            data, box = self.data, self.box
        if dev is None:
            dev = DeviceSystem(self._device)
            dev.set_atoms(data["x"], data["y"], data["z"], box.box, box.origin, box.boundary)
        self.dev = dev
        M, real_max = dev.build_neighbor(self.rc, self.max_neigh)
        if self.max_neigh is not None and real_max > self.max_neigh:
            raise ValueError(
                f"max_neigh={self.max_neigh} is too small: at least one "
                f"atom has {real_max} neighbors within rc={self.rc}. "
                f"Re-run with max_neigh>={real_max} (or omit max_neigh "
                "to let mdapy size the buffer automatically)."
            )
        self._host = [None, None, None]
        if fetch:
            self._host = list(dev.fetch_neighbor())

    def _get(self, k):
        if self._host[k] is None:
            if self.dev is None:
                raise AttributeError("call compute() first")
            want = [False, False, False]
            want[k] = True
            self._host[k] = self.dev.fetch_neighbor(*want)[k]
        return self._host[k]

    @property
    def verlet_list(self) -> np.ndarray:
        return self._get(0)

    @property
    def distance_list(self) -> np.ndarray:
        return self._get(1)

    @property
    def neighbor_number(self) -> np.ndarray:
        return self._get(2)
