#!/usr/bin/env python
"""Benchmark input: a periodic Voronoi polycrystal generated on the GPU (BASELINE configs[3]).

Follows the recipe of the reference's ``CreatePolycrystal`` (src/mdapy/create_polycrystal.py:117-137 seeds and
Euler angles from ``np.random.default_rng(seed)``: ``seed_position = rng.random((G,3))*L``,
``theta = rng.uniform(-180,180,(G,3))``, R = Rx Ry Rz, 286-297; grain g keeps the rotated lattice points whose
nearest periodic seed is g -- the Voronoi-cell plane test of src/polycrystal.cpp:88-103 -- and atoms of
different grains closer than ``overlap`` are thinned keeping the lower index, src/neighbor.cpp:465-476).
It is an INPUT GENERATOR for tools/bench_configs_dist.py, not the parity-checked GPU port of that class (a
"next" row of SURVEY.md 8f.2): boundary atoms are decided by distance comparison instead of voro++ face
planes, so positions agree with the reference's construction but membership at exact ties may differ.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

FCC = np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.0], [0.0, 0.5, 0.5], [0.5, 0.0, 0.5]])


def _rot(theta_deg, axis):
    t = np.deg2rad(theta_deg)
    c, s = np.cos(t), np.sin(t)
    x, y, z = axis
    return np.array([[c + x * x * (1 - c), x * y * (1 - c) - z * s, x * z * (1 - c) + y * s],
                     [y * x * (1 - c) + z * s, c + y * y * (1 - c), y * z * (1 - c) - x * s],
                     [z * x * (1 - c) - y * s, z * y * (1 - c) + x * s, c + z * z * (1 - c)]])


def make_polycrystal(L: float, n_grains: int, a: float, seed: int, device, overlap: float = 2.0, reach: float = 1.9,
                     chunk: int = 1 << 21, basis=FCC):
    """x, y, z (torch f64 on ``device``) of an L^3 periodic polycrystal with ``n_grains`` grains of lattice
    constant ``a``.  ``reach``: lattice points are generated within reach * (L^3/G)^(1/3) of each seed."""
    import torch

    rng = np.random.default_rng(seed)
    seeds = rng.random((n_grains, 3)) * L
    theta = rng.uniform(-180, 180, (n_grains, 3))
    S = torch.tensor(seeds, dtype=torch.float64, device=device)
    B = torch.tensor(np.asarray(basis) * a, dtype=torch.float64, device=device)
    rg = min(reach * (L ** 3 / n_grains) ** (1.0 / 3.0), 0.5 * L * np.sqrt(3.0))
    m = int(np.ceil(rg / a)) + 1
    ax = torch.arange(-m, m + 1, dtype=torch.float64, device=device) * a
    out = []
    for g in range(n_grains):
        R = torch.tensor(_rot(theta[g, 0], (1.0, 0, 0)) @ _rot(theta[g, 1], (0, 1.0, 0)) @ _rot(theta[g, 2], (0, 0, 1.0)),
                         dtype=torch.float64, device=device)
        # lattice planes in chunks of x to bound memory
        step = max(1, chunk // ((2 * m + 1) ** 2 * B.shape[0]))
        for i0 in range(0, 2 * m + 1, step):
            gx = ax[i0:i0 + step]
            P = torch.stack(torch.meshgrid(gx, ax, ax, indexing="ij"), dim=-1).reshape(-1, 1, 3) + B.view(1, -1, 3)
            P = P.reshape(-1, 3)
            P = P[(P * P).sum(dim=1) <= rg * rg]
            if P.numel() == 0:
                continue
            P = P @ R.T + S[g]
            P = P - torch.floor(P / L) * L                       # wrap into the box
            # nearest periodic seed must be g (fp32 is enough to decide membership)
            d = (P.view(-1, 1, 3) - S.view(1, -1, 3)).to(torch.float32)
            d = d - torch.round(d / L) * L
            d2 = (d * d).sum(dim=2)
            keep = torch.argmin(d2, dim=1) == g
            out.append(P[keep])
    pos = torch.cat(out)
    x, y, z = (pos[:, k].contiguous() for k in range(3))
    if overlap > 0:
        keep = thin_overlaps(x, y, z, L, overlap, device)
        x, y, z = x[keep].contiguous(), y[keep].contiguous(), z[keep].contiguous()
    return x, y, z, seeds


def thin_overlaps(x, y, z, L, overlap, device):
    """Mask of atoms to keep: atom j goes when a lower-index atom i lies within ``overlap`` (pairs found with
    the library's own cut-off list).  One sweep in index order like the reference: an atom already removed does
    not remove others."""
    import torch

    from mdapy_b200.device import DeviceSystem

    idx = device.index if device.index is not None else 0
    ds = DeviceSystem(idx)
    box = np.diag([L, L, L]).astype(float)
    ds.set_atoms_device(x, y, z, box, np.zeros(3), np.array([1, 1, 1], np.int32))
    M, mx = ds.build_neighbor(float(overlap))
    v, _, n = ds.fetch_neighbor(want_dist=False)
    del ds
    N = x.shape[0]
    alive = np.ones(N, bool)
    close = np.nonzero(n > 0)[0]
    for i in close:                                   # few atoms (grain-boundary overlaps only), index order
        if not alive[i]:
            continue
        for j in v[i, : n[i]]:
            if j > i:
                alive[j] = False
    return torch.tensor(alive, device=x.device)


if __name__ == "__main__":
    import torch

    L = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    G = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    x, y, z, seeds = make_polycrystal(L, G, 4.05, 1, torch.device("cuda", 0))
    n_ideal = 4 * (L / 4.05) ** 3
    print(f"atoms {x.numel()}  ({x.numel() / n_ideal:.4f} of the perfect-crystal count)")
