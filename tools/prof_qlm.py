"""One Steinhardt q4,q6 call on a 2 M-atom thermal FCC frame, meant to run under ncu (kernel filter k_qlm)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
import numpy as np, torch
from profile_all import lattice, FCC
from mdapy_b200.device import DeviceSystem
dev = torch.device("cuda", 0)
(x, y, z), box = lattice(FCC, 4.05, 80, 0.12, 4, dev)
ds = DeviceSystem(0)
ds.set_atoms_device(x, y, z, box, np.zeros(3), np.array([1, 1, 1], np.int32))
rc = 0.85 * 4.05
ds.build_neighbor(rc)
ds.steinhardt([4, 6], rc=rc, average=False, fetch=False)
torch.cuda.synchronize()
