"""Polyhedral template matching, mirroring ``mdapy.polyhedral_template_matching.PolyhedralTemplateMatching``
(src/mdapy/polyhedral_template_matching.py:16-167).  ``output`` (N, 8): structure type (0 other, 1 fcc,
2 hcp, 3 bcc, 4 ico, 5 sc, 6 cubic diamond, 7 hexagonal diamond, 8 graphene), alloy ordering (1 pure, 2 L1_0,
3 L1_2 Cu, 4 L1_2 Au, 5 B2, 6 SiC, 7 BN), rmsd, interatomic distance, quaternion w, x, y, z;
``ptm_indices`` (N, 18): the atom and its matched neighbours (first 18 points of the matched environment).

All eight reference structures are built ("default" = fcc-hcp-bcc-ico); the diamond and graphene ones read
the ranked neighbour lists of the first-shell atoms (ptm_multishell.cpp:94-184)."""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import tool_function as tool
from .box import Box
from .device import LIST_KNN, DeviceSystem
from .frame import Frame
from .knn import NearestNeighbor

_STRUCTURES = ["fcc", "hcp", "bcc", "ico", "sc", "dcub", "dhex", "graphene", "all", "default"]


class PolyhedralTemplateMatching:
    def __init__(self, structure: str, data, box: Box, rmsd_threshold: float = 0.1,
                 verlet_list: Optional[np.ndarray] = None, dev: Optional[DeviceSystem] = None, device: int = 0):
        self.structure = structure
        self.data = Frame.from_any(data)
        self.box = box
        self.rmsd_threshold = rmsd_threshold
        self.verlet_list = verlet_list
        self._dev = dev
        self._device = device
        for i in self.structure.split("-"):
            assert i in _STRUCTURES, (
                'Structure should in ["fcc", "hcp", "bcc", "ico", "sc","dcub", "dhex", "graphene", "all", "default"].'
            )

    def compute(self) -> None:
        N = self.data.shape[0]
        if sum(self.box.boundary) == 0 and N <= 18:
            self.output = np.zeros((N, 7))          # reference quirk kept (Appendix D.7)
            self.ptm_indices = np.zeros((N, 18), np.int32)
            return
        box, data, dev = self.box, self.data, self._dev
        safe_L = 15
        if dev is None and self.verlet_list is None:
            repeat = np.ceil(safe_L / self.box.get_thickness()).astype(int)
            for i in range(3):
                if self.box.boundary[i] == 0:
                    repeat[i] = 1
            if sum(repeat) != 3:
                data, box = tool._replicate_pos(data, box, *repeat)
                # per-atom types follow the replication (the reference tiles the frame it was given)
                for col in ("type", "element"):
                    if col in self.data.columns:
                        data = data.with_columns(**{col: np.tile(np.asarray(self.data[col]), int(np.prod(repeat)))})
            knn = NearestNeighbor(data, box, 18, device=self._device)
            knn.compute(fetch=False)
            dev = knn.dev
            if hasattr(knn, "_enlarge_data"):
                data, box = knn._enlarge_data, knn._enlarge_box
        elif dev is None:
            dev = DeviceSystem(self._device)
            dev.set_atoms(data["x"], data["y"], data["z"], box.box, box.origin, box.boundary)
            dev.put_neighbor(self.verlet_list, kind=LIST_KNN)
        if "type" in data.columns:
            type_list = np.asarray(data["type"]).astype(np.int32)
        elif "element" in data.columns:
            el = np.asarray(data["element"])
            ele2type = {j: i + 1 for i, j in enumerate(sorted(set(el.tolist())))}
            type_list = np.array([ele2type[e] for e in el.tolist()], np.int32)
        else:
            type_list = np.ones(data.shape[0], np.int32)
        self.output, self.ptm_indices = dev.ptm(self.structure, self.rmsd_threshold, type_list)
