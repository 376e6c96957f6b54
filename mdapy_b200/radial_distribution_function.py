"""Radial distribution function, mirroring ``mdapy.radial_distribution_function.RadialDistributionFunction``
(src/mdapy/radial_distribution_function.py:20-279): pair counts on the GPU (list kernels or the streaming
kernel), normalisation in NumPy exactly as the reference (147-211).  Attributes: ``r``, ``g_total``,
``g_partial`` keyed by (label_a, label_b), ``elements``, ``Ntype``, ``type_list``."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from .box import Box
from .device import LIST_CUTOFF, DeviceSystem


class RadialDistributionFunction:
    def __init__(self, rc: float, nbin: int, box: Box, verlet_list=None, distance_list=None, neighbor_number=None,
                 type_list=None, streaming: bool = False, x=None, y=None, z=None,
                 dev: Optional[DeviceSystem] = None, device: int = 0):
        self.rc = float(rc)
        self.nbin = int(nbin)
        self.box = box
        self.vol = self.box.volume
        self.streaming = bool(streaming)
        self._dev = dev
        self._device = device
        if self.streaming:
            if dev is None and (x is None or y is None or z is None):
                raise ValueError("streaming=True requires x, y, z position arrays.")
            if x is not None:
                self._x = np.ascontiguousarray(x, dtype=np.float64)
                self._y = np.ascontiguousarray(y, dtype=np.float64)
                self._z = np.ascontiguousarray(z, dtype=np.float64)
                assert self._x.shape == self._y.shape == self._z.shape, "x, y, z must have the same shape"
                self.N = int(self._x.shape[0])
            else:
                self.N = dev.N
            self.verlet_list = self.distance_list = self.neighbor_number = None
        else:
            if dev is None and (verlet_list is None or distance_list is None or neighbor_number is None):
                raise ValueError("streaming=False requires verlet_list, distance_list, neighbor_number.")
            self.verlet_list = verlet_list
            self.distance_list = distance_list
            self.neighbor_number = neighbor_number
            self.N = int(verlet_list.shape[0]) if verlet_list is not None else dev.n_rows
        raw = np.zeros(self.N, dtype=np.int32) if type_list is None else np.asarray(type_list)
        unique_sorted = sorted(set(raw.tolist()))
        self.elements: List[Any] = list(unique_sorted)
        self.Ntype = len(self.elements)
        label_to_idx = {label: i for i, label in enumerate(self.elements)}
        self.type_list = np.array([label_to_idx[v] for v in raw.tolist()], dtype=np.int32)

    def compute(self) -> None:
        edges = np.linspace(0, self.rc, self.nbin + 1)
        const = (4.0 * np.pi / 3.0 * (edges[1:] ** 3 - edges[:-1] ** 3)) / self.vol
        self.r = (edges[1:] + edges[:-1]) / 2
        dev = self._dev
        if dev is None:
            dev = DeviceSystem(self._device)
            b = self.box
            if self.streaming:
                dev.set_atoms(self._x, self._y, self._z, b.box, b.origin, b.boundary)
            else:
                # list kernels never touch coordinates; a 1-atom placeholder keeps the handle valid
                z = np.zeros(self.N)
                dev.set_atoms(z, z, z, b.box, b.origin, b.boundary)
                dev.put_neighbor(self.verlet_list, self.distance_list, self.neighbor_number, rc=self.rc,
                                 kind=LIST_CUTOFF)
        counts = np.zeros((self.Ntype, self.Ntype, self.nbin), dtype=np.float64)
        if self.streaming:
            counts = dev.rdf_counts(self.rc, self.nbin, self.type_list, self.Ntype, streaming=True)
        elif self.Ntype > 1:
            counts = dev.rdf_counts(self.rc, self.nbin, self.type_list, self.Ntype, streaming=False)
        else:
            counts[0, 0] = dev.rdf_counts(self.rc, self.nbin, None, 1, streaming=False)
        self.counts = counts
        number_per_type = np.bincount(self.type_list, minlength=self.Ntype)
        total = np.zeros(self.nbin, dtype=np.float64)
        for a in range(self.Ntype):
            for b in range(self.Ntype):
                total += counts[a, b]
        self.g_total = total / const / self.N**2
        self.g_partial: Dict[Tuple[Any, Any], np.ndarray] = {}
        for a in range(self.Ntype):
            n_a = number_per_type[a]
            for b in range(a, self.Ntype):
                n_b = number_per_type[b]
                raw = counts[a, b] if a == b else counts[a, b] + counts[b, a]
                if n_a > 0 and n_b > 0:
                    g_ab = raw / (n_a * n_b) / const
                    if a != b:
                        g_ab *= 0.5
                else:
                    g_ab = np.zeros_like(self.r)
                self.g_partial[(self.elements[a], self.elements[b])] = g_ab
