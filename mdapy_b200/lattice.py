"""Perfect-crystal supercells for the benchmark configurations, mirroring the slice of
``mdapy.build_crystal`` the hot path needs (src/mdapy/build_lattice.py:657-907, bases 30-61;
replication order of src/repeat_cell.cpp:41-59: cell-major, iz fastest,
``pos = basis*a + (ix*a1 + iy*a2 + iz*a3)``).  Miller-rotated and multi-species cells are
outside the hot path."""
from __future__ import annotations

import numpy as np

from .box import Box
from .tool_function import repeat_cell

_BASES = {
    "sc": np.array([[0.0, 0.0, 0.0]]),
    "fcc": np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.0], [0.0, 0.5, 0.5], [0.5, 0.0, 0.5]]),
    "bcc": np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]]),
    "diamond": np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.0], [0.0, 0.5, 0.5], [0.5, 0.0, 0.5],
                         [0.25, 0.25, 0.25], [0.75, 0.75, 0.25], [0.25, 0.75, 0.75], [0.75, 0.25, 0.75]]),
}


def crystal_positions(structure: str, a: float, nx: int = 1, ny: int = 1, nz: int = 1):
    """(pos[N,3], box[3,3]) of a cubic-cell crystal in standard orientation."""
    s = structure.lower()
    if s not in _BASES:
        raise ValueError(f"structure {structure!r} is not available here; supported: {sorted(_BASES)}")
    old_box = a * np.eye(3)
    old_pos = _BASES[s] @ old_box
    new_pos = repeat_cell(old_box, old_pos, nx, ny, nz)
    new_box = old_box * np.array([nx, ny, nz]).reshape((3, 1))
    return new_pos, new_box


def build_crystal(name, structure: str, a: float, miller1=None, miller2=None, miller3=None,
                  nx: int = 1, ny: int = 1, nz: int = 1, c=None):
    if miller1 is not None or miller2 is not None or miller3 is not None:
        raise NotImplementedError("Miller-rotated cells are outside the hot path (SURVEY.md 2.2)")
    from .system import System

    pos, box = crystal_positions(structure, a, nx, ny, nz)
    el = name if isinstance(name, str) else name[0]
    data = {"x": pos[:, 0].copy(), "y": pos[:, 1].copy(), "z": pos[:, 2].copy(),
            "element": np.full(pos.shape[0], el, dtype=object)}
    return System(data=data, box=Box(box))
