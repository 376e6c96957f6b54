"""k-nearest build time vs the cell-width factor (MDB_KNN_CELL) on rattled BCC / FCC frames."""
import sys, os, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
import numpy as np, torch
from profile_all import lattice, FCC, BCC
from mdapy_b200.device import DeviceSystem
dev = torch.device("cuda", 0)
o, bnd = np.zeros(3), np.array([1, 1, 1], np.int32)
for name, basis, a, n in (("bcc", BCC, 2.8665, 160), ("fcc", FCC, 3.615, 120)):
    (x, y, z), box = lattice(basis, a, n, 0.05, 2, dev)
    ds = DeviceSystem(0)
    for k in (12, 18):
        best = 1e9
        for _ in range(3):
            ds.set_atoms_device(x, y, z, box, o, bnd)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            ds.build_knn(k)
            ds.synchronize(); torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        print(os.environ.get("MDB_KNN_CELL", "default"), name, x.numel(), "k", k, f"{best*1e3:.1f} ms", f"{x.numel()/best/1e6:.0f} M atoms/s", flush=True)
    del ds
