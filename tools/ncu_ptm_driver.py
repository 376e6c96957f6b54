"""Driver for an ncu capture of the PTM kernels: 1.02 M rattled BCC atoms, kNN(18) + PTM fcc-hcp-bcc."""
import sys
import numpy as np
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import helpers as H
from mdapy_b200.device import DeviceSystem
p, b = H.bcc(2.8665, 80)
pos = H.rattle(p, 0.05, 0)
x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
ds = DeviceSystem(0)
ds.set_atoms(x, y, z, b, np.zeros(3), np.array([1, 1, 1], np.int32))
ds.build_knn(18)
ds.ptm("fcc-hcp-bcc", fetch=False)
ds.ptm("fcc-hcp-bcc", fetch=False)
ds.synchronize()
