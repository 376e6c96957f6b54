#!/usr/bin/env python
"""Per-kernel throughput of the other BASELINE configs on one GPU (device-resident, CUDA-synchronised
wall clock around each library call, best of 3), next to the reference C++ on the host cores on a
bounded sample.  Writes gpurun_out/r2_config_throughput_1gpu.json (copied to profiles/).  Not the headline bench (that is bench.py)."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import helpers as H  # noqa: E402
from mdapy_b200.device import DeviceSystem  # noqa: E402
from oracle import checker as K  # noqa: E402


def timed(fn, reps=3):
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


def lattice_dev(basis, a, n, sigma, seed, dev):
    b = torch.tensor(basis, dtype=torch.float64, device=dev) * a
    g = torch.arange(n, dtype=torch.float64, device=dev) * a
    nb = b.shape[0]
    gen = torch.Generator(device=dev).manual_seed(seed)
    cols = []
    for d in range(3):
        shape = [1, 1, 1, 1]
        shape[d] = n
        c = (g.view(shape) + b[:, d].view(1, 1, 1, nb)).expand(n, n, n, nb).reshape(-1)
        if sigma > 0:
            c = c + torch.randn(c.shape, dtype=torch.float64, device=dev, generator=gen) * sigma
        cols.append(c.contiguous())
    return cols, np.diag([n * a] * 3).astype(float)


def main():
    dev = torch.device("cuda", 0)
    cores = os.cpu_count()
    o, bnd = np.zeros(3), np.array([1, 1, 1], np.int32)
    report = {"gpu": torch.cuda.get_device_name(0), "host_cores": cores, "checker": K.KIND, "rows": []}

    def row(config, kernel, n_atoms, sec, cpu=None):
        r = {"config": config, "kernel": kernel, "atoms": n_atoms, "gpu_ms": sec * 1e3, "gpu_atoms_per_s": n_atoms / sec}
        if cpu:
            r.update(cpu_atoms_per_s=cpu[0], cpu_sample_atoms=cpu[1], speedup=n_atoms / sec / cpu[0])
        report["rows"].append(r)
        print(json.dumps(r), flush=True)

    # ---- config 2: 10 M FCC Cu, neighbour + CNA + CSP(12) -------------------------------------------
    (x, y, z), box = lattice_dev(H.FCC, 3.615, 136, 0.02, 1, dev)
    N = x.numel()
    rc = 3.615 * 0.8536
    ds = DeviceSystem(0)
    ds.set_atoms_device(x, y, z, box, o, bnd)

    def nb():
        ds.set_atoms_device(x, y, z, box, o, bnd)
        ds.build_neighbor(rc)

    t_nb = timed(nb)
    t_cna = timed(lambda: ds.fcna(rc, fetch=False))
    ds.set_atoms_device(x, y, z, box, o, bnd)
    ds.build_neighbor(4.2)                       # >= 12 neighbours everywhere: CSP via sorted cut-off list
    t_sort = timed(lambda: ds.sort_neighbor(12), reps=1)
    t_csp = timed(lambda: ds.csp(12, fetch=False))
    # CPU sample: 1 M atoms
    ps, bs = H.fcc(3.615, 63)
    ps = H.rattle(ps, 0.02, 1)
    xs, ys, zs = (np.ascontiguousarray(ps[:, k]) for k in range(3))
    t0 = time.perf_counter(); v, d, n = K.build_neighbor_auto(xs, ys, zs, bs, o, bnd, rc); c_nb = time.perf_counter() - t0
    t0 = time.perf_counter(); K.fcna(xs, ys, zs, bs, o, bnd, v, n, rc); c_cna = time.perf_counter() - t0
    t0 = time.perf_counter(); ki, kd = K.knn(xs, ys, zs, bs, o, bnd, 12); c_knn12 = time.perf_counter() - t0
    t0 = time.perf_counter(); K.csp(xs, ys, zs, bs, o, bnd, ki, 12); c_csp = time.perf_counter() - t0
    ns = xs.shape[0]
    row("C2 10M FCC Cu", "binning+neighbour(auto M)", N, t_nb, (ns / c_nb, ns))
    row("C2 10M FCC Cu", "fixed CNA", N, t_cna, (ns / c_cna, ns))
    row("C2 10M FCC Cu", "sort 12 of rc=4.2 list", N, t_sort)
    row("C2 10M FCC Cu", "CSP(12)", N, t_csp, (ns / c_csp, ns))
    t_knn = timed(lambda: ds.build_knn(12))
    row("C2 10M FCC Cu", "kNN(12)", N, t_knn, (ns / c_knn12, ns))
    del ds, x, y, z
    torch.cuda.empty_cache()

    # ---- config 3: BCC Fe, PTM + Ackland-Jones (kNN 18 / 14) ----------------------------------------
    nb3 = int(os.environ.get("C3_N", "200"))     # 200^3*2 = 16 M; 292 -> 49.8 M
    (x, y, z), box = lattice_dev(H.BCC, 2.8665, nb3, 0.05, 2, dev)
    N = x.numel()
    ds = DeviceSystem(0)
    ds.set_atoms_device(x, y, z, box, o, bnd)
    t_k18 = timed(lambda: ds.build_knn(18), reps=2)
    t_ptm = timed(lambda: ds.ptm("fcc-hcp-bcc", 0.1, None, fetch=False), reps=2)
    t_k14 = timed(lambda: ds.build_knn(14), reps=2)
    t_aja = timed(lambda: ds.aja(fetch=False))
    ps, bs = H.bcc(2.8665, 50)                   # 250 k atoms for the CPU (PTM pre-ordering is serial there)
    ps = H.rattle(ps, 0.05, 2)
    xs, ys, zs = (np.ascontiguousarray(ps[:, k]) for k in range(3))
    ns = xs.shape[0]
    t0 = time.perf_counter(); ki, kd = K.knn(xs, ys, zs, bs, o, bnd, 18); c_k18 = time.perf_counter() - t0
    c_ptm = None
    if K.KIND == "reference":
        t0 = time.perf_counter(); K.ptm("fcc-hcp-bcc", xs, ys, zs, bs, o, bnd, ki, np.ones(ns, np.int32), 0.1); c_ptm = time.perf_counter() - t0
    t0 = time.perf_counter(); ki, kd = K.knn(xs, ys, zs, bs, o, bnd, 14); c_k14 = time.perf_counter() - t0
    t0 = time.perf_counter(); K.aja(xs, ys, zs, bs, o, bnd, ki, kd); c_aja = time.perf_counter() - t0
    row(f"C3 {N/1e6:.1f}M BCC Fe", "kNN(18)", N, t_k18, (ns / c_k18, ns))
    row(f"C3 {N/1e6:.1f}M BCC Fe", "PTM fcc-hcp-bcc", N, t_ptm, (ns / c_ptm, ns) if c_ptm else None)
    row(f"C3 {N/1e6:.1f}M BCC Fe", "kNN(14)", N, t_k14, (ns / c_k14, ns))
    row(f"C3 {N/1e6:.1f}M BCC Fe", "Ackland-Jones", N, t_aja, (ns / c_aja, ns))
    del ds, x, y, z
    torch.cuda.empty_cache()

    # ---- config 4 stand-in: 20 M thermal FCC Al, Steinhardt q4/q6 + RDF(500 bins, rc = 6) -----------
    (x, y, z), box = lattice_dev(H.FCC, 4.05, 171, 0.12, 3, dev)
    N = x.numel()
    ds = DeviceSystem(0)
    ds.set_atoms_device(x, y, z, box, o, bnd)
    rq = 0.85 * 4.05
    ds.build_neighbor(rq)
    t_q = timed(lambda: ds.steinhardt([4, 6], rc=rq, fetch=False))
    t_qa = timed(lambda: ds.steinhardt([4, 6], rc=rq, average=True, fetch=False))
    types = np.zeros(N, np.int32)
    t_rs = timed(lambda: ds.rdf_counts(6.0, 500, types, 1, streaming=True), reps=2)
    ds.set_atoms_device(x, y, z, box, o, bnd)
    t_l6 = timed(lambda: (ds.set_atoms_device(x, y, z, box, o, bnd), ds.build_neighbor(6.0)), reps=2)   # the first call allocates ~30 GB
    t_rl = timed(lambda: ds.rdf_counts(6.0, 500, None, 1, streaming=False), reps=2)
    ps, bs = H.fcc(4.05, 50)
    ps = H.rattle(ps, 0.12, 3)
    xs, ys, zs = (np.ascontiguousarray(ps[:, k]) for k in range(3))
    ns = xs.shape[0]
    v, d, n = K.build_neighbor_auto(xs, ys, zs, bs, o, bnd, rq)
    t0 = time.perf_counter(); K.get_sq(xs, ys, zs, bs, o, bnd, v, d, n, [4, 6], rc=rq); c_q = time.perf_counter() - t0
    t0 = time.perf_counter(); K.rdf_streaming(xs, ys, zs, np.zeros(ns, np.int32), 1, bs, o, bnd, 6.0, 500); c_rs = time.perf_counter() - t0
    row("C4* 20M thermal FCC Al", "Steinhardt q4,q6 (list rc=0.85a)", N, t_q, (ns / c_q, ns))
    row("C4* 20M thermal FCC Al", "Steinhardt q4,q6 averaged", N, t_qa)
    row("C4* 20M thermal FCC Al", "RDF streaming rc=6 500 bins", N, t_rs, (ns / c_rs, ns))
    row("C4* 20M thermal FCC Al", "neighbour list rc=6 (auto M)", N, t_l6)
    row("C4* 20M thermal FCC Al", "RDF from list rc=6 500 bins", N, t_rl)
    out = ROOT / "gpurun_out" / "r2_config_throughput_1gpu.json"
    out.parent.mkdir(exist_ok=True)
    out.write_text(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
