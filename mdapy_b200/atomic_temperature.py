"""Local atomic temperature, mirroring ``mdapy.atomic_temperature.AtomicTemperature``
(src/mdapy/atomic_temperature.py:11-118; kernel: src/atomic_temperature.cpp:7-112).

Masses come from an ``amass`` column.  The reference can also look masses up from an ``element`` column through
its periodic-table data module, which is outside this path: pass ``amass`` explicitly."""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _lib as L
from .device import DeviceSystem
from .frame import Frame


class AtomicTemperature:
    def __init__(self, data, verlet_list: Optional[np.ndarray] = None, distance_list: Optional[np.ndarray] = None,
                 rc: float = 5.0, factor: float = 1.0, dev: Optional[DeviceSystem] = None) -> None:
        self.data = Frame.from_any(data)
        self.verlet_list = verlet_list
        self.distance_list = distance_list
        self.rc = rc
        self.factor = factor
        self._dev = dev

    def compute(self) -> None:
        for i in ["vx", "vy", "vz"]:
            assert i in self.data.columns, "No velocity information."
        if "amass" not in self.data.columns:
            raise ValueError("No atomic mass." if "element" not in self.data.columns else
                             "No atomic mass: add an 'amass' column (the element -> mass table is outside this path).")
        amass = np.asarray(self.data["amass"], np.float64)
        vx, vy, vz = (np.asarray(self.data[c], np.float64) * 1e3 * self.factor for c in ("vx", "vy", "vz"))
        if self._dev is not None:
            self.T = self._dev.atomic_temperature(vx, vy, vz, amass, float(self.rc))
            return
        v, d = L.i32(self.verlet_list), L.f64(self.distance_list)
        N, M = v.shape
        self.T = np.zeros(N, float)
        L.check(L.lib().mdb_compute_temp(L.iptr(v), N, M, L.dptr(d), L.dptr(L.f64(vx)), L.dptr(L.f64(vy)), L.dptr(L.f64(vz)),
                                         L.dptr(L.f64(amass)), L.dptr(self.T), float(self.rc), 1))
