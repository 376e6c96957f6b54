"""Voronoi polycrystal builder on the device, mirroring the metal-only path of
``mdapy.create_polycrystal.CreatePolycrystal`` (src/mdapy/create_polycrystal.py:88-147 seeds / angles,
259-314 grain generation, 583-680 replication, 684-840 overlap removal + wrap):

* seeds ``rng.random((G, 3)) * L`` and Euler angles ``rng.uniform(-180, 180, (G, 3))`` from
  ``np.random.default_rng(randomseed)``, R = Rx Ry Rz (Rodrigues, degrees);
* the unit cell is replicated ``ceil(r_max / thickness)`` times (r_max: largest cavity radius of the periodic
  Voronoi cells) -- ``repeat_cell`` on the device, once;
* every grain: ``transform_and_filter`` (rotate about the block centre, move to the seed, keep what lies inside
  the cell's face planes) on the device-resident block;
* ``filter_overlap_atom`` (drop the higher index of every pair closer than ``overlap``) and wrap on the device.

The reference takes cell faces and cavity radii from voro++; here they come from ``scipy.spatial.Voronoi`` on
the 27 periodic copies of the seeds (same cells; plane coefficients are the face bisectors).  Graphene
decoration of grain boundaries (``add_graphene``) is outside the hot path."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib as L
from . import builders as B
from .box import Box
from .lattice import _BASES


def _rodrigues(theta_deg: float, axis) -> np.ndarray:
    """create_polycrystal.py:150-201."""
    x, y, z = np.asarray(axis, float) / np.linalg.norm(axis)
    t = np.deg2rad(theta_deg)
    c, s = np.cos(t), np.sin(t)
    return np.array([[c + x * x * (1 - c), x * y * (1 - c) - z * s, x * z * (1 - c) + y * s],
                     [y * x * (1 - c) + z * s, c + y * y * (1 - c), y * z * (1 - c) - x * s],
                     [z * x * (1 - c) - y * s, z * y * (1 - c) + x * s, c + z * z * (1 - c)]])


class CreatePolycrystal:
    def __init__(self, structure: str, a: float, box, seed_number: int, seed_position: Optional[np.ndarray] = None,
                 theta_list: Optional[np.ndarray] = None, randomseed: Optional[int] = None, metal_overlap_dis: float = 2.0,
                 need_rotation: bool = True, add_graphene: bool = False, device: int = 0):
        if add_graphene:
            raise NotImplementedError("graphene-decorated grain boundaries are outside the hot path")
        self.structure, self.a = structure.lower(), float(a)
        if self.structure not in _BASES:
            raise ValueError(f"structure {structure!r} is not available here; supported: {sorted(_BASES)}")
        self.box = box if isinstance(box, Box) else Box(np.eye(3) * float(box) if np.isscalar(box) else box)
        lengths = np.diag(self.box.box).astype(float)
        assert np.allclose(self.box.box, np.diag(lengths)), "orthogonal boxes only"
        self.lengths = lengths
        self.seed_number = int(seed_number)
        self.randomseed = np.random.randint(0, 10000000) if randomseed is None else int(randomseed)
        self.rng = np.random.default_rng(self.randomseed)
        self.seed_position = (self.rng.random((self.seed_number, 3)) * lengths if seed_position is None
                              else np.asarray(seed_position, float))
        self.need_rotation = need_rotation
        self.theta_list = (self.rng.uniform(-180, 180, (self.seed_number, 3)) if theta_list is None
                           else np.asarray(theta_list, float))
        self.metal_overlap_dis = float(metal_overlap_dis)
        self.device = int(device)

    # ---- periodic Voronoi cells of the seeds: face planes and cavity radius per grain
    def _cells(self):
        from scipy.spatial import Voronoi

        G, Lb = self.seed_number, self.lengths
        shifts = np.array([[i, j, k] for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1)], float) * Lb
        order = np.argsort(np.abs(shifts).sum(axis=1), kind="stable")      # the unshifted copy first
        shifts = shifts[order]
        pts = (self.seed_position[None, :, :] + shifts[:, None, :]).reshape(-1, 3)   # pts[0:G] = the seeds themselves
        vor = Voronoi(pts)
        planes = [[] for _ in range(G)]
        for (p, q) in vor.ridge_points:
            for g, h in ((p, q), (q, p)):
                if g < G:
                    s, t = pts[g], pts[h]
                    n = t - s
                    planes[g].append([2 * n[0], 2 * n[1], 2 * n[2], -(t @ t - s @ s)])   # |r - s|^2 < |r - t|^2
        radius = np.zeros(G)
        for g in range(G):
            verts = vor.vertices[[v for v in vor.regions[vor.point_region[g]] if v >= 0]]
            # voro++ keeps vertices at double scale: the reference's cavity_radius = sqrt(max_radius_squared())
            # (src/voronoi.cpp:66) is TWICE the seed-to-farthest-vertex distance, which is what makes the
            # replicated block (edge >= r_max) cover the whole cell around its centre
            radius[g] = 2.0 * np.sqrt(((verts - pts[g]) ** 2).sum(axis=1).max())
        return [np.asarray(p, float) for p in planes], radius

    def compute(self, verbose: bool = False):
        from .system import System

        planes, radius = self._cells()
        r_max = float(radius.max())
        reps = np.ceil(r_max / self.a).astype(int) * np.ones(3, int)     # cubic unit cell: thickness = a
        block = B.device_lattice(self.structure, self.a, int(reps[0]), int(reps[1]), int(reps[2]), device=self.device)
        # centre of the block = mean position (create_polycrystal.py:301); same value for every grain
        bx, by, bz = B.fetch_positions(block)
        centre = L.f64(np.array([bx.mean(), by.mean(), bz.mean()]))
        self.block_centre = centre.copy()   # (the reference takes polars' column mean; any mean differs in the last bits)
        del bx, by, bz
        out = np.empty((block.N, 3), np.float64)
        pos_list, grain_list = [], []
        for g in range(self.seed_number):
            if self.need_rotation:
                th = self.theta_list[g]
                R = _rodrigues(th[0], (1.0, 0, 0)) @ _rodrigues(th[1], (0, 1.0, 0)) @ _rodrigues(th[2], (0, 0, 1.0))
            else:
                R = _rodrigues(0, (1.0, 0, 0))
            R, tgt, pl = L.f64(R), L.f64(self.seed_position[g]), L.f64(planes[g])
            n = C.c_int(0)
            L.check(block._lib.mdb_system_transform_and_filter(block._h, L.dptr(R), L.dptr(centre), L.dptr(tgt), L.dptr(pl),
                                                               pl.shape[0], L.dptr(out), C.byref(n)))
            pos_list.append(out[: n.value].copy())
            grain_list.append(np.full(n.value, g + 1, np.int32))
            if verbose:
                print(f"  grain {g + 1}/{self.seed_number}: {n.value} atoms")
        block.close()
        pos = np.concatenate(pos_list)
        grain = np.concatenate(grain_list)
        pos += self.box.origin[None, :]
        x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
        if self.metal_overlap_dis > 0:
            keep = B.filter_overlap_atom(x, y, z, self.box.box, self.box.origin, self.box.boundary, self.metal_overlap_dis)
            x, y, z, grain = x[keep], y[keep], z[keep], grain[keep]
        system = System(data={"x": x, "y": y, "z": z, "grain_id": grain, "type": np.ones(x.shape[0], np.int32)},
                        box=self.box, device=self.device)
        system.wrap_pos()
        return system
