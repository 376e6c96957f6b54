"""Warren-Cowley short-range-order parameter, mirroring
``mdapy.warren_cowley_parameter.WarrenCowleyParameter`` (src/mdapy/warren_cowley_parameter.py:15-112;
kernel: src/warren_cowley_parameter.cpp:9-80).  Plotting is outside the hot path."""
from __future__ import annotations

from typing import Optional

import numpy as np

from .device import DeviceSystem
from .frame import Frame


class WarrenCowleyParameter:
    def __init__(self, verlet_list: Optional[np.ndarray], neighbor_number: Optional[np.ndarray], data,
                 dev: Optional[DeviceSystem] = None, device: int = 0) -> None:
        self.verlet_list = verlet_list
        self.neighbor_number = neighbor_number
        self.data = Frame.from_any(data)
        # warren_cowley_parameter.py:77-94: 0-based types from 'element' (sorted unique symbols) or 'type'
        if "element" in self.data.columns:
            el = np.asarray(self.data["element"])
            uniq = sorted(set(el.tolist()))
            self.ele2type = {e: i for i, e in enumerate(uniq)}
            self.type_list = np.array([self.ele2type[e] for e in el.tolist()], np.int32)
            self.Ntype = len(uniq)
        else:
            assert "type" in self.data.columns, "data must contain an 'element' or 'type' column."
            t = np.asarray(self.data["type"]).astype(np.int32)
            self.type_list = np.ascontiguousarray(t - 1, dtype=np.int32)
            self.Ntype = int(np.unique(t).size)
            assert int(self.type_list.max()) + 1 == self.Ntype
        self._dev = dev
        self._device = device

    def compute(self) -> None:
        dev = self._dev
        if dev is None:
            # the list alone is enough; positions are not read by this kernel
            from . import _lib as L

            v, n = L.i32(self.verlet_list), L.i32(self.neighbor_number)
            self.WCP = np.zeros((self.Ntype, self.Ntype), float)
            L.check(L.lib().mdb_get_wcp(L.iptr(v), v.shape[0], v.shape[1], L.iptr(n), L.iptr(self.type_list),
                                        self.Ntype, L.dptr(self.WCP), 1))
            return
        self.WCP = dev.wcp(self.type_list, self.Ntype)
