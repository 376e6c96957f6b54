#!/usr/bin/env python
"""tools/group_probe.py -- end-to-end time of System(devices=[...]).cal_common_neighbor_analysis(rc) on the
BASELINE configs[4] frame (99.6 M-atom FCC Al) from ONE unpartitioned host array, for page-locked and for
pageable input, with the device group's phase times.

    python tools/group_probe.py [--n 292] [--devices 0,1,...] [--reps 3] [--sigma 0.0]

Prints one JSON line per (devices, input kind).  Labels are checked (perfect lattice: all fcc; rattled: equal
to the single-GPU labels)."""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import mdapy_b200 as mp  # noqa: E402
from mdapy_b200 import _lib as L  # noqa: E402


def lattice(n, a):
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]]) * a
    g = np.arange(n, dtype=np.float64) * a
    x = np.empty((n, n, n, 4))
    y = np.empty_like(x)
    z = np.empty_like(x)
    for k in range(4):
        x[..., k] = g[:, None, None] + basis[k, 0]
        y[..., k] = g[None, :, None] + basis[k, 1]
        z[..., k] = g[None, None, :] + basis[k, 2]
    return x.reshape(-1), y.reshape(-1), z.reshape(-1), np.diag([n * a] * 3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=292)
    ap.add_argument("--devices", default="")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--sigma", type=float, default=0.0)
    args = ap.parse_args()
    import ctypes as C

    ndev = C.c_int(0)
    L.check(L.lib().mdb_device_count(C.byref(ndev)))
    sets = [[int(v) for v in args.devices.split(",")]] if args.devices else \
        [list(range(k)) for k in (1, 2, 4, 8) if k <= ndev.value]
    a = 4.05
    rc = a * 0.8536
    x, y, z, box = lattice(args.n, a)
    if args.sigma > 0:
        rng = np.random.default_rng(0)
        for v in (x, y, z):
            v += rng.normal(0.0, args.sigma, v.shape)
    N = x.shape[0]
    pinned = [L.result_empty(N, np.float64) for _ in range(3)]
    for dst, src in zip(pinned, (x, y, z)):
        dst[:] = src
    ref = None
    for devs in sets:
        for kind, (hx, hy, hz) in (("pinned", pinned), ("pageable", (x, y, z))):
            times, phases = [], None
            for rep in range(args.reps + 1):            # first repetition is the warm-up
                t0 = time.perf_counter()
                kw = {"devices": devs} if len(devs) > 1 else {"device": devs[0]}
                s = mp.System(data={"x": hx, "y": hy, "z": hz}, box=mp.Box(box), **kw)
                s.cal_common_neighbor_analysis(rc)
                lab = np.asarray(s.data["cna"])
                dt = (time.perf_counter() - t0) * 1e3
                if rep:
                    times.append(dt)
                if s._group is not None:
                    phases = s._group.last_times()
                    used = s._group.members_used
                else:
                    used = 1
                if rep == 0:
                    if args.sigma == 0:
                        assert int(lab.min()) == 1 and int(lab.max()) == 1, "perfect FCC: all labels 1"
                    elif ref is None:
                        ref = lab.copy()
                    else:
                        assert np.array_equal(lab, ref), "labels differ from the first configuration"
                del s, lab
            ms = float(np.median(times))
            print(json.dumps({"devices": devs, "members_used": used, "input": kind, "atoms": N,
                              "ms": round(ms, 2), "ms_each": [round(t, 2) for t in times],
                              "atoms_per_s": N / ms * 1e3, "phases_ms": phases}), flush=True)


if __name__ == "__main__":
    main()
