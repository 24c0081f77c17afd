import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package()
scene, desc = pkg.example("pile", (40, 12345, 2.2, 41))
b = pkg.Batch(scene, n_worlds=1, device=0, coloured=True)
b.set_scene_forces(desc)
for f in range(100): b.step()
pairs, colours = b.pair_levels(0)
print('pairs', len(pairs), 'max colour', colours.max(), 'hist', np.bincount(colours)[:70])
c0=b.counters(); b.step(); b.sync(); c1=b.counters()
print({k:c1[k]-c0[k] for k in c0})
