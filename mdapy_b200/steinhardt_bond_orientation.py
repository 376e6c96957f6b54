"""Steinhardt bond orientation, mirroring ``mdapy.steinhardt_bond_orientation.SteinhardtBondOrientation``
(src/mdapy/steinhardt_bond_orientation.py:16-302): q_l, optional w_l / w_l-hat, neighbour averaging and
the solid/liquid bond criterion.  Results: ``qnarray``, ``qlm_r``, ``qlm_i`` (+ ``solidliquid``, ``nbond``)."""
from __future__ import annotations

from typing import Optional

import numpy as np

from .box import Box
from .device import LIST_CUTOFF, LIST_KNN, DeviceSystem
from .frame import Frame


class SteinhardtBondOrientation:
    def __init__(self, box: Box, data, llist, nnn: int, rc: float, average: bool, use_voronoi: bool,
                 use_weight: bool, weight: Optional[np.ndarray], verlet_list: Optional[np.ndarray] = None,
                 distance_list: Optional[np.ndarray] = None, neighbor_number: Optional[np.ndarray] = None,
                 wl: bool = False, wlhat: bool = False, identify_liquid: bool = False, threshold: float = 0.7,
                 n_bond: int = 7, dev: Optional[DeviceSystem] = None, device: int = 0):
        self.box = box
        self.data = Frame.from_any(data)
        self.llist = np.asarray(llist, int)
        self.nnn = int(nnn)
        self.rc = rc
        self.average = average
        self.use_voronoi = use_voronoi
        self.use_weight = use_weight
        self.weight = weight
        self.verlet_list = verlet_list
        self.distance_list = distance_list
        self.neighbor_number = neighbor_number
        self.wl = wl
        self.wlhat = wlhat
        self.identify_liquid = identify_liquid
        self.threshold = threshold
        self.n_bond = n_bond
        self._dev = dev
        self._device = device

    def compute(self) -> None:
        if self.identify_liquid:
            assert 6 in self.llist
            assert self.threshold > 0
            assert self.n_bond > 0
        if not self.use_voronoi and self.nnn <= 0:
            assert self.rc > 0
        dev = self._dev
        if dev is None:
            dev = DeviceSystem(self._device)
            d, b = self.data, self.box
            dev.set_atoms(d["x"], d["y"], d["z"], b.box, b.origin, b.boundary)
            dev.put_neighbor(self.verlet_list, self.distance_list, self.neighbor_number,
                             rc=self.rc if self.rc and self.rc > 0 else -1.0,
                             kind=LIST_KNN if self.nnn > 0 else LIST_CUTOFF)
        weight = None
        if self.use_weight:
            assert self.weight is not None and self.weight.shape == (dev.n_rows, dev.M)
            weight = self.weight
        self.qnarray, self.qlm_r, self.qlm_i = dev.steinhardt(
            self.llist, nnn=self.nnn, rc=self.rc, average=self.average, wl=self.wl, wlhat=self.wlhat,
            use_voronoi=self.use_voronoi, weight=weight, fetch_qlm=True)
        if self.identify_liquid:
            q6index = int(np.where(self.llist == 6)[0][0])
            self.solidliquid, self.nbond = dev.solid_liquid(q6index, float(self.threshold), int(self.n_bond),
                                                            nnn=self.nnn, rc=self.rc, use_voronoi=self.use_voronoi)
