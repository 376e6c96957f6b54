"""Bond-length / bond-angle distributions and the angular distribution function, mirroring
``mdapy.bond_analysis.BondAnalysis`` (src/mdapy/bond_analysis.py:17-160) and
``mdapy.angular_distribution_function.AngularDistributionFunction``
(src/mdapy/angular_distribution_function.py:17-168); kernels: src/bond_analysis.cpp:7-240.  Plotting is
outside the hot path."""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np

from .box import Box
from .device import LIST_CUTOFF, DeviceSystem
from .frame import Frame


def _device_with_list(data: Frame, box: Box, verlet_list, distance_list, neighbor_number, rc, device: int):
    dev = DeviceSystem(device)
    dev.set_atoms(data["x"], data["y"], data["z"], box.box, box.origin, box.boundary)
    dev.put_neighbor(verlet_list, distance_list, neighbor_number, rc=float(rc), kind=LIST_CUTOFF)
    return dev


class BondAnalysis:
    def __init__(self, data, box: Box, rc: float, nbin: int, verlet_list: Optional[np.ndarray] = None,
                 distance_list: Optional[np.ndarray] = None, neighbor_number: Optional[np.ndarray] = None,
                 dev: Optional[DeviceSystem] = None, device: int = 0):
        self.data = Frame.from_any(data)
        self.box = box
        self.rc = rc
        self.nbin = nbin
        self.verlet_list = verlet_list
        self.distance_list = distance_list
        self.neighbor_number = neighbor_number
        self._dev = dev
        self._device = device

    def compute(self):
        dev = self._dev or _device_with_list(self.data, self.box, self.verlet_list, self.distance_list,
                                             self.neighbor_number, self.rc, self._device)
        self.bond_length_distribution, self.bond_angle_distribution = dev.bond_analysis(float(self.rc), int(self.nbin))
        r = np.linspace(0, self.rc, self.nbin + 1)
        self.r_length = (r[1:] + r[:-1]) / 2
        r = np.linspace(0, 180.0, self.nbin + 1)
        self.r_angle = (r[1:] + r[:-1]) / 2


class AngularDistributionFunction:
    def __init__(self, data, box: Box, rc_dict: Dict[str, List[float]], nbin: int,
                 verlet_list: Optional[np.ndarray] = None, distance_list: Optional[np.ndarray] = None,
                 neighbor_number: Optional[np.ndarray] = None, dev: Optional[DeviceSystem] = None, device: int = 0):
        self.data = Frame.from_any(data)
        assert "element" in self.data.columns
        self.box = box
        self.ele_unique = sorted(set(np.asarray(self.data["element"]).tolist()))
        pair_list = []
        for key in rc_dict.keys():
            a, b, c = key.split("-")
            assert a in self.ele_unique
            assert b in self.ele_unique
            assert c in self.ele_unique
            pair_list.append([self.ele_unique.index(a), self.ele_unique.index(b), self.ele_unique.index(c)])
        self.pair_list = np.array(pair_list, np.int32)
        self.rc_list = np.array(list(rc_dict.values()), float)
        assert self.rc_list.shape[1] == 4, "rc should be a list of 4 floats."
        self.nbin = nbin
        self.verlet_list = verlet_list
        self.distance_list = distance_list
        self.neighbor_number = neighbor_number
        self._dev = dev
        self._device = device

    def compute(self):
        ele2type = {j: i for i, j in enumerate(self.ele_unique)}
        type_list = np.array([ele2type[e] for e in np.asarray(self.data["element"]).tolist()], np.int32)
        dev = self._dev or _device_with_list(self.data, self.box, self.verlet_list, self.distance_list,
                                             self.neighbor_number, float(self.rc_list.max()), self._device)
        self.bond_angle_distribution = dev.adf(self.rc_list, self.pair_list, type_list, int(self.nbin))
        r = np.linspace(0, 180.0, self.nbin + 1)
        self.r_angle = (r[1:] + r[:-1]) / 2
