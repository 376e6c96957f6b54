import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import helpers as H
import test_gpu_slab as T
a = 4.05
p, box = H.fcc(a, 40)
pos = H.rattle(p, 0.08, 7)
rng = np.random.default_rng(2)
hot = pos[:, 1] > 0.6 * box[1, 1]
pos[hot] += rng.normal(0, 0.4, (int(hot.sum()), 3))
T._full_and_slabs(pos, box, [1,1,1], 0.85*a, 2)
print("ok")
