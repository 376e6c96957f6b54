import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
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
