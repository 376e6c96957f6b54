"""Voronoi tessellation, mirroring ``mdapy.voronoi.Voronoi`` (src/mdapy/voronoi.py:32-400): cell volume, face
count and cavity radius per atom, and Voronoi neighbours with face areas.  The cells are built on the GPU
(csrc/voronoi.cu) instead of by voro++: orthogonal boxes with any mix of periodic and open boundaries (open axes end
at the box faces), and triclinic boxes the way the reference treats them -- every axis periodic, open axes tripled
first (voronoi.py:148-152; voro++'s container_triclinic is periodic)."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import tool_function as tool
from .box import Box
from .device import DeviceSystem
from .frame import Frame


class Voronoi:
    def __init__(self, box: Box, data, dev: Optional[DeviceSystem] = None, device: int = 0):
        self.box = box
        self.data = Frame.from_any(data)
        self._dev = dev
        self._device = device

    def _device_for(self, box: Box, data: Frame, reuse: bool, nopbc: bool = False) -> DeviceSystem:
        cell, boundary = box.box, box.boundary
        if box.triclinic and not nopbc and sum(box.boundary) < 3:
            # voronoi.py:148-152: the open axes of a triclinic cell are tripled and the container is periodic.  (The
            # reference also rotates the cell into LAMMPS form first; the cells do not depend on the orientation.)
            cell = np.array(box.box, float)
            for i in range(3):
                if box.boundary[i] == 0:
                    cell[i] *= 3
            boundary = np.ones(3, np.int32)
            reuse = False
        if reuse and self._dev is not None:
            return self._dev
        dev = DeviceSystem(self._device)
        dev.set_atoms(data["x"], data["y"], data["z"], cell, box.origin, boundary)
        return dev

    def get_neighbor(self, a_face_area_threshold: float = -1.0,
                     r_face_area_threshold: float = -1.0) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        """voronoi.py:71-180.  Fewer than 50 atoms in a periodic box: the frame is replicated first, exactly as the
        reference does, and the arrays describe the replicated frame (``_enlarge_data`` / ``_enlarge_box``)."""
        repeat = [1, 1, 1]
        N = self.data.shape[0]
        if N < 50:
            if sum(self.box.boundary) > 0:
                while np.prod(repeat) * N < 50:
                    for i in range(3):
                        if self.box.boundary[i] == 1:
                            repeat[i] += 1
            else:
                assert N > 1, "system with all free boundary must has at least 2 atoms."
        data, box = self.data, self.box
        if sum(repeat) != 3:
            self._enlarge_data, self._enlarge_box = tool._replicate_pos(data, box, *repeat)
            data, box = self._enlarge_data, self._enlarge_box
        dev = self._device_for(box, data, reuse=sum(repeat) == 3)
        return dev.voronoi_neighbor(a_face_area_threshold, r_face_area_threshold)

    def get_volume(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """voronoi.py:262-330 -> (volume, neighbor_number, cavity_radius)."""
        dev = self._device_for(self.box, self.data, reuse=True)
        return dev.voronoi_volume()
