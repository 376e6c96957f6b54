#!/usr/bin/env python
"""tools/voronoi_probe.py -- Voronoi cells at benchmark size: time of mdb_system_voronoi_volume /
mdb_system_voronoi_neighbor on an n^3 x 4 rattled FCC frame, with the reference's voro++ (oracle/_ref, all host
threads) timed and compared on a sub-frame.

    python tools/voronoi_probe.py [--n 136] [--ref-n 40] [--sigma 0.05]"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import helpers as H  # noqa: E402
from mdapy_b200.device import DeviceSystem  # noqa: E402
from oracle import checker as K  # noqa: E402


def frame(n, a, sigma, seed):
    pos, box = H.fcc(a, n)
    pos = pos + np.random.default_rng(seed).normal(0.0, sigma, pos.shape)
    return tuple(np.ascontiguousarray(pos[:, k]) for k in range(3)) + (box,)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=136)
    ap.add_argument("--ref-n", type=int, default=40)
    ap.add_argument("--sigma", type=float, default=0.05)
    args = ap.parse_args()
    a = 3.615
    O3, PBC = np.zeros(3), np.array([1, 1, 1], np.int32)
    out = {}
    # parity + reference time on the sub-frame
    x, y, z, box = frame(args.ref_n, a, args.sigma, 1)
    t0 = time.perf_counter()
    rvol, rnn, rrad = K.voronoi_volume(x, y, z, box, O3, PBC)
    t_ref = time.perf_counter() - t0
    ds = DeviceSystem(0)
    ds.set_atoms(x, y, z, box, O3, PBC)
    vol, nn, rad = ds.voronoi_volume()
    out["parity"] = {"atoms": int(x.shape[0]), "faces_equal": bool(np.array_equal(nn, rnn)),
                     "volume_max_rel": float(np.abs(vol / rvol - 1).max()),
                     "radius_max_rel": float(np.abs(rad / rrad - 1).max())}
    out["reference"] = {"atoms": int(x.shape[0]), "s": t_ref, "atoms_per_s": x.shape[0] / t_ref,
                        "threads": K.num_threads()}
    ds.close()
    # device time at size
    x, y, z, box = frame(args.n, a, args.sigma, 0)
    ds = DeviceSystem(0)
    ds.set_atoms(x, y, z, box, O3, PBC)
    ds.voronoi_volume()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        vol, nn, rad = ds.voronoi_volume()
        ts.append(time.perf_counter() - t0)
    N = x.shape[0]
    out["device_volume"] = {"atoms": N, "s": min(ts), "atoms_per_s": N / min(ts), "faces_hist": np.bincount(nn).tolist(),
                            "volume_sum_rel_err": float(abs(vol.sum() / np.prod(np.diag(box)) - 1))}
    t0 = time.perf_counter()
    v, d, ar, n2 = ds.voronoi_neighbor(-1.0, 0.01)
    out["device_neighbor"] = {"s": time.perf_counter() - t0, "M": int(v.shape[1])}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
