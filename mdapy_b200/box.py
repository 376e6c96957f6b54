"""Simulation cell, mirroring the reference's ``mdapy.box.Box`` (src/mdapy/box.py:6-509).

Same constructor forms (scalar, 3-vector, 3x3, 4x3 with origin row, 3x4 OVITO
layout, copy), same properties and the same helper arithmetic
(``get_thickness`` box.py:465-481, ``check_small_box`` box.py:483-502) so that
replication decisions agree with the reference.
"""
from __future__ import annotations

import numpy as np


class Box:
    def __init__(self, box, boundary=None, origin=None) -> None:
        if isinstance(box, Box):
            self._box = box.box.copy()
            self._origin = box.origin.copy()
            self._boundary = box.boundary.copy()
        else:
            self._box, self._origin = self._parse(box, origin)
            self.set_boundary(boundary)
        self._refresh()

    # -- parsing ---------------------------------------------------------
    @staticmethod
    def _parse_origin(origin):
        if origin is None:
            return np.zeros(3, np.float64)
        if isinstance(origin, (list, tuple, np.ndarray)):
            origin = np.array(origin, np.float64)
            if origin.shape != (3,):
                raise ValueError(f"Origin must be a 3-element array, got shape {origin.shape}")
            return origin
        raise TypeError(f"Invalid origin type: {type(origin)}")

    @classmethod
    def _parse(cls, box, origin):
        if isinstance(box, (int, float)):
            box = np.eye(3, dtype=np.float64) * box
        elif isinstance(box, (list, tuple, np.ndarray)):
            box = np.array(box, np.float64)
            if box.shape == (3,):
                box = np.diag(box)
            elif box.shape == (3, 3):
                pass
            elif box.shape == (4, 3):
                origin = np.array(box[-1])
                box = np.array(box[:-1])
            elif box.shape == (3, 4):
                origin = np.array(box[:, -1])
                box = np.array(box[:, :-1])
            else:
                raise ValueError(f"Invalid box shape: {box.shape}")
        else:
            raise TypeError(f"Invalid box type: {type(box)}")
        return np.ascontiguousarray(box), cls._parse_origin(origin)

    def _refresh(self):
        b = self._box
        tri = False
        for i in range(3):
            for j in range(3):
                if i != j and abs(b[i, j]) > 1e-10:
                    tri = True
        if np.any(np.diag(b) < 0):
            tri = True
        self._triclinic = tri
        self._inverse = np.linalg.inv(b)
        self._volume = float(np.linalg.det(b))

    # -- setters ---------------------------------------------------------
    def set_boundary(self, boundary=None) -> None:
        if boundary is None:
            boundary = np.array([1, 1, 1], np.int32)
        elif isinstance(boundary, (list, tuple, np.ndarray)):
            boundary = np.array(boundary, np.int32)
            if boundary.shape != (3,):
                raise ValueError(f"Boundary must be a 3-element array, got shape {boundary.shape}")
            boundary = np.where(boundary != 0, 1, 0).astype(np.int32)
        else:
            raise TypeError(f"Invalid boundary type: {type(boundary)}")
        self._boundary = boundary

    def set_box(self, box) -> None:
        self._box, _ = self._parse(box, self._origin)
        self._refresh()

    def set_origin(self, origin) -> None:
        self._origin = self._parse_origin(origin)

    # -- properties ------------------------------------------------------
    @property
    def box(self) -> np.ndarray:
        return self._box

    @property
    def origin(self) -> np.ndarray:
        return self._origin

    @property
    def boundary(self) -> np.ndarray:
        return self._boundary

    @property
    def triclinic(self) -> bool:
        return self._triclinic

    @property
    def inverse_box(self) -> np.ndarray:
        return self._inverse

    @property
    def volume(self) -> float:
        return self._volume

    # -- helpers ---------------------------------------------------------
    def pbc(self, rij: np.ndarray) -> np.ndarray:
        """Minimum-image displacement (box.py:443-463)."""
        rij = np.asarray(rij, float) @ self.inverse_box
        for i in range(3):
            if self.boundary[i] == 1:
                rij[i] -= np.floor(rij[i] + 0.5)
        return rij @ self.box

    def get_thickness(self) -> np.ndarray:
        b = self.box
        return np.array(
            [
                self.volume / np.linalg.norm(np.cross(b[1], b[2])),
                self.volume / np.linalg.norm(np.cross(b[0], b[2])),
                self.volume / np.linalg.norm(np.cross(b[0], b[1])),
            ],
            dtype=np.float64,
        )

    def check_small_box(self, rc: float) -> np.ndarray:
        thickness = self.get_thickness()
        repeat = np.ones(3, dtype=np.int32)
        for i in range(3):
            if self.boundary[i] == 1 and thickness[i] < 2 * rc:
                repeat[i] = int(np.ceil(2.0 * rc / thickness[i]))
        return repeat

    def __repr__(self) -> str:
        return f"Box(box={self.box.tolist()}, boundary={self.boundary.tolist()}, origin={self.origin.tolist()})"
