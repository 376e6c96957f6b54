#!/usr/bin/env python
"""tools/parity_at_scale.py -- parity against the reference at the sizes BASELINE.json names.

TEST INFRASTRUCTURE (imports oracle/): compares the CUDA path with the reference's own C++
(oracle/_ref, compiled unmodified from src/neighbor.cpp, src/cna.cpp, src/centro_symmetry_parameter.cpp)
on the same seeded inputs, full arrays, element for element:

  c2        BASELINE configs[1]: 10,061,824-atom FCC Cu (136^3 x 4), rattled sigma = 0.05 (seed 0):
            Neighbor(rc = 0.8536 a, automatic width) rows / distances / counts, FixedCNA labels,
            sort_verlet_by_distance(12) + get_csp(12)                                   -- all bit-equal
  c2hot     the same lattice at sigma = 0.20 (seed 1): varied counts, non-trivial labels, padded rows
  c5        BASELINE configs[4]: 99,588,352-atom FCC Al (292^3 x 4), rattled sigma = 0.05 (seed 2),
            build_neighbor(max_neigh = 16) + FixedCNA (neighbor.cpp:351, cna.cpp:429): 64-bit row offsets,
            400 k-CTA grids, chunked label copies
  c5auto    the same frame through the automatic width (sampled estimate + compaction), rows vs c5's

    python tools/parity_at_scale.py c2 c2hot            # ~1 min
    python tools/parity_at_scale.py c5 c5auto           # ~6 min, ~60 GB of host memory

Prints one line per check and a final JSON summary; exit code 1 on any mismatch.
"""
from __future__ import annotations

import json
import os
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

O3 = np.zeros(3)
PBC = np.array([1, 1, 1], np.int32)
RESULTS = []


def check(name, ok, detail=""):
    RESULTS.append({"check": name, "ok": bool(ok), "detail": detail})
    print(f"[{'ok' if ok else 'MISMATCH'}] {name} {detail}", flush=True)


def equal_chunked(a, b, rows=1 << 22):
    """np.array_equal on big 2-D arrays without a second full-size temporary; bit pattern for f64."""
    if a.shape != b.shape:
        return False
    if a.dtype == np.float64:
        a, b = a.view(np.int64), b.view(np.int64)
    for s in range(0, a.shape[0], rows):
        if not np.array_equal(a[s:s + rows], b[s:s + rows]):
            return False
    return True


def frame(basis, a, n, sigma, seed):
    pos, box = H.lattice(basis, a, n, n, n)
    if sigma > 0:
        rng = np.random.default_rng(seed)
        for s in range(0, pos.shape[0], 1 << 22):       # chunked: no (N,3) normal() temporary at 100 M atoms
            pos[s:s + (1 << 22)] += rng.normal(0.0, sigma, pos[s:s + (1 << 22)].shape)
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    del pos
    return x, y, z, box


def timed(label, fn):
    t0 = time.perf_counter()
    out = fn()
    print(f"    {label}: {time.perf_counter() - t0:.1f} s", flush=True)
    return out


def case_c2(tag, sigma, seed, with_csp):
    a = 3.615
    rc = a * 0.8536
    x, y, z, box = frame(H.FCC, a, 136, sigma, seed)
    N = x.shape[0]
    print(f"== {tag}: {N} atoms FCC Cu, sigma={sigma}, rc={rc}", flush=True)
    rv, rd, rn = timed("reference build_neighbor_without_max_neigh", lambda: K.build_neighbor_auto(x, y, z, box, O3, PBC, rc))
    rp = timed("reference FixedCNA", lambda: K.fcna(x, y, z, box, O3, PBC, rv, rn, rc))
    ds = DeviceSystem(0)
    ds.set_atoms(x, y, z, box, O3, PBC)
    M, mx = ds.build_neighbor(rc, None)
    v, d, nn = ds.fetch_neighbor()
    check(f"{tag}.width", (M, mx) == (rv.shape[1], int(rn.max())), f"M={M} max={mx} ref={rv.shape[1]}")
    check(f"{tag}.neighbor_number", np.array_equal(nn, rn), f"min={int(nn.min())} max={int(nn.max())}")
    check(f"{tag}.verlet_list (row order included)", equal_chunked(v, rv))
    check(f"{tag}.distance_list (bit pattern)", equal_chunked(d, rd))
    p = ds.fcna(rc)
    hist = np.bincount(p, minlength=5).tolist()
    check(f"{tag}.cna", np.array_equal(p, rp), f"labels other/fcc/hcp/bcc/ico = {hist}")
    pf, used = ds.fused_cna(rc)
    check(f"{tag}.fused neighbour+CNA (no list)", used and np.array_equal(pf, rp))
    if with_csp and int(rn.min()) >= 12:
        # system.py:1986-2003: the cached cut-off list is sorted (12 smallest first), then get_csp
        K.sort_verlet_by_distance(rv, rd, 12)
        rc_ = timed("reference get_csp", lambda: K.csp(x, y, z, box, O3, PBC, rv, 12))
        ds.sort_neighbor(12)
        c = ds.csp(12)
        v2, d2, _ = ds.fetch_neighbor()
        check(f"{tag}.sorted rows", equal_chunked(v2, rv) and equal_chunked(d2, rd))
        check(f"{tag}.csp (bit pattern)", np.array_equal(c.view(np.int64), rc_.view(np.int64)),
              f"max={float(c.max()):.6f}")
    ds.close()


def case_c5(auto_too):
    a = 4.05
    rc = a * 0.8536
    x, y, z, box = frame(H.FCC, a, 292, 0.05, 2)
    N = x.shape[0]
    print(f"== c5: {N} atoms FCC Al, sigma=0.05, rc={rc}, max_neigh=16", flush=True)
    rv, rd, rn = timed("reference build_neighbor(max_neigh=16)", lambda: K.build_neighbor(x, y, z, box, O3, PBC, rc, 16))
    rp = timed("reference FixedCNA", lambda: K.fcna(x, y, z, box, O3, PBC, rv, rn, rc))
    ds = DeviceSystem(0)
    ds.set_atoms(x, y, z, box, O3, PBC)
    M, mx = ds.build_neighbor(rc, 16)
    v, d, nn = timed("device fetch (19 GB)", lambda: ds.fetch_neighbor())
    check("c5.neighbor_number", np.array_equal(nn, rn), f"min={int(nn.min())} max={int(nn.max())} rows*M={N * 16}")
    check("c5.verlet_list (row order included)", equal_chunked(v, rv))
    check("c5.distance_list (bit pattern)", equal_chunked(d, rd))
    p = ds.fcna(rc)
    check("c5.cna", np.array_equal(p, rp), f"labels = {np.bincount(p, minlength=5).tolist()}")
    pf, used = ds.fused_cna(rc)
    check("c5.fused neighbour+CNA (no list)", used and np.array_equal(pf, rp))
    if auto_too:
        del v, d
        M2, mx2 = ds.build_neighbor(rc, None)
        v, d, nn2 = ds.fetch_neighbor()
        mref = int(rn.max())
        check("c5auto.width", (M2, mx2) == (mref, mref), f"M={M2}")
        check("c5auto.rows", np.array_equal(nn2, rn) and equal_chunked(v, rv[:, :M2]) and equal_chunked(d, rd[:, :M2]))
        p2 = ds.fcna(rc)
        check("c5auto.cna", np.array_equal(p2, rp))
    ds.close()


def main():
    which = sys.argv[1:] or ["c2", "c2hot"]
    print(f"checker: oracle.{K.KIND}, host threads {os.cpu_count()}", flush=True)
    if "c2" in which:
        case_c2("c2", 0.05, 0, True)
    if "c2hot" in which:
        case_c2("c2hot", 0.20, 1, False)
    if "c5" in which or "c5auto" in which:
        case_c5("c5auto" in which)
    bad = [r for r in RESULTS if not r["ok"]]
    print(json.dumps({"checks": len(RESULTS), "mismatches": len(bad), "checker": K.KIND,
                      "failed": [r["check"] for r in bad]}), flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
