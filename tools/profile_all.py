#!/usr/bin/env python
"""One launch of every kernel on the path at a moderate size, meant to run under ncu:

    ncu --set full --clock-control none --import-source on --kernel-name regex:"^k_|::k_" -o gpurun_out/all_kernels \
        python tools/profile_all.py

Frames: rattled FCC Al (n = 100 -> 4.0 M atoms) for the cut-off consumers, rattled BCC Fe (n = 110 -> 2.66 M atoms)
for the k-nearest consumers.  tools/ncu_table.py turns the report into profiles/<name>.md."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from mdapy_b200.device import DeviceSystem  # noqa: E402

FCC = [[0.0, 0.0, 0.0], [0.5, 0.5, 0.0], [0.0, 0.5, 0.5], [0.5, 0.0, 0.5]]
BCC = [[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]]


def lattice(basis, a, n, sigma, seed, dev):
    b = torch.tensor(basis, dtype=torch.float64, device=dev) * a
    g = torch.arange(n, dtype=torch.float64, device=dev) * a
    nb = b.shape[0]
    gen = torch.Generator(device=dev).manual_seed(seed)
    cols = []
    for d in range(3):
        shape = [1, 1, 1, 1]
        shape[d] = n
        c = (g.view(shape) + b[:, d].view(1, 1, 1, nb)).expand(n, n, n, nb).reshape(-1)
        c = c + torch.randn(c.shape, dtype=torch.float64, device=dev, generator=gen) * sigma
        cols.append(c.contiguous())
    return cols, np.diag([n * a] * 3).astype(float)


def main():
    dev = torch.device("cuda", 0)
    o, bnd = np.zeros(3), np.array([1, 1, 1], np.int32)
    # ---- cut-off consumers
    (x, y, z), box = lattice(FCC, 4.05, int(sys.argv[1]) if len(sys.argv) > 1 else 100, 0.05, 1, dev)
    rc = 4.05 * 0.8536
    ds = DeviceSystem(0)
    ds.set_atoms_device(x, y, z, box, o, bnd)
    ds.build_neighbor(rc)                      # binning kernels, count-only sample, tile kernel (+ compaction)
    ds.fcna(rc, fetch=False)
    ds.steinhardt([4, 6], rc=rc, average=True, wl=True, fetch=False)
    ds.solid_liquid(1, 0.7, 7, rc=rc)
    ds.cnp(rc, fetch=False)
    ds.cluster(rc)
    ds.structure_entropy(rc, 0.2, False, float(np.linalg.det(box)))
    N = x.numel()
    types = np.zeros(N, np.int32)
    ds.rdf_counts(rc, 200, type_list=types, ntype=1)
    ds.rdf_counts(rc, 200, type_list=types, ntype=1, streaming=True)
    ds.wcp(types, 1)
    ds.bond_analysis(rc, 90)
    ds.adf(np.array([[0.0, rc, 0.0, rc]]), np.array([[0, 0, 0]], np.int32), types, 90)
    vel = np.random.default_rng(0).standard_normal((3, N))
    ds.atomic_temperature(vel[0], vel[1], vel[2], np.full(N, 26.98), rc)
    ds.set_atoms_device(x, y, z, box, o, bnd)
    ds.build_neighbor(4.4)                     # >= 14 neighbours: sort + CSP + AJA from the cut-off list
    ds.sort_neighbor(14)
    ds.csp(12, fetch=False)
    ds.aja(fetch=False)
    ds.acna(fetch=False)
    del ds, x, y, z
    # ---- k-nearest consumers
    (x, y, z), box = lattice(BCC, 2.8665, int(sys.argv[2]) if len(sys.argv) > 2 else 110, 0.05, 2, dev)
    ds = DeviceSystem(0)
    ds.set_atoms_device(x, y, z, box, o, bnd)
    ds.build_knn(18)
    ds.ptm("fcc-hcp-bcc", 0.1, fetch=False)
    ds.build_knn(4)
    ds.ids(fetch=False)
    torch.cuda.synchronize()
    print("profile_all done", N, x.numel())


if __name__ == "__main__":
    main()
