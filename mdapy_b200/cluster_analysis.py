"""Cluster analysis, mirroring ``mdapy.cluster_analysis.ClusterAnalysis``
(src/mdapy/cluster_analysis.py:14-132; kernels: src/cluster.cpp:9-150).

Cluster ids are 1-based and numbered by the smallest atom index of each cluster, exactly like the reference's
serial flood from ascending seeds (identical for symmetric bond lists, which every cut-off list is)."""
from __future__ import annotations

from typing import Dict, Optional, Union

import numpy as np

from . import _lib as L
from .device import DeviceSystem


def type_pair_table(rc: Dict[str, float]):
    """cluster_analysis.py:75-89: 'a-b' keys -> (type1, type2, r) rows, both orders for unlike pairs."""
    t1, t2, r = [], [], []
    for key, value in rc.items():
        left, right = key.split("-")
        t1.append(left)
        t2.append(right)
        r.append(value)
        if left != right:
            t1.append(right)
            t2.append(left)
            r.append(value)
    return np.array(t1, np.int32), np.array(t2, np.int32), np.array(r, float)


class ClusterAnalysis:
    def __init__(self, rc: Union[float, int, Dict[str, float]], verlet_list: Optional[np.ndarray] = None,
                 distance_list: Optional[np.ndarray] = None, neighbor_number: Optional[np.ndarray] = None,
                 type_list: Optional[np.ndarray] = None, dev: Optional[DeviceSystem] = None):
        self.rc = rc
        if isinstance(rc, (float, int, np.integer, np.floating)):
            self.max_rc = self.rc
        elif isinstance(rc, dict):
            assert type_list is not None, "Need type_list for multi cutoff mode."
            self.max_rc = max([i for i in self.rc.values()])
        else:
            raise TypeError("rc should be a positive number, or a dict like {'1-1':1.5, '1-2':1.3}")
        self.verlet_list = verlet_list.copy() if (isinstance(rc, dict) and verlet_list is not None) else verlet_list
        self.distance_list = distance_list
        self.neighbor_number = neighbor_number
        self.type_list = type_list
        self._dev = dev

    def _filter_verlet(self):
        """Filter the host neighbour list according to type-dependent cut-offs (cluster.cpp:114-150)."""
        t1, t2, r = type_pair_table(self.rc)
        v, d, n = self.verlet_list, L.f64(self.distance_list), L.i32(self.neighbor_number)
        assert v.dtype == np.int32 and v.flags.c_contiguous
        L.check(L.lib().mdb_filter_by_type(L.iptr(v), v.shape[0], v.shape[1], L.dptr(d), L.iptr(n),
                                           L.iptr(L.i32(self.type_list)), L.iptr(t1), L.iptr(t2), L.dptr(r),
                                           int(t1.shape[0]), 1))

    def compute(self):
        if self._dev is not None:
            dev = self._dev
            if isinstance(self.rc, dict):
                t1, t2, r = type_pair_table(self.rc)
                self.particleClusters, self.cluster_number = dev.cluster(0.0, self.type_list, t1, t2, r)
            else:
                self.particleClusters, self.cluster_number = dev.cluster(float(self.max_rc))
            return
        if isinstance(self.rc, dict):
            self._filter_verlet()
        v, n = L.i32(self.verlet_list), L.i32(self.neighbor_number)
        N, M = v.shape
        self.particleClusters = np.full(N, -1, dtype=np.int32)
        import ctypes as C

        cnt = C.c_int(0)
        if isinstance(self.rc, dict):
            L.check(L.lib().mdb_get_cluster_by_bond(L.iptr(v), N, M, L.iptr(n), L.iptr(self.particleClusters),
                                                    C.byref(cnt)))
        else:
            d = L.f64(self.distance_list)
            L.check(L.lib().mdb_get_cluster(L.iptr(v), N, M, L.dptr(d), L.iptr(n), float(self.max_rc),
                                            L.iptr(self.particleClusters), C.byref(cnt)))
        self.cluster_number = cnt.value
