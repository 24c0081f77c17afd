import os, sys, subprocess, numpy as np
ROOT='/root/repo'
lib=os.path.join(ROOT,'raw-physics_b200','librawphys_b200_f32.so')
out='/tmp/f32.npz'
r=subprocess.run([sys.executable, os.path.join(ROOT,'tests','f32_worker.py'), out], capture_output=True, text=True, env=dict(os.environ, RAWPHYS_B200_LIB=lib))
print(r.stdout[-500:], r.stderr[-1500:])
z=np.load(out); G=np.load(os.path.join(ROOT,'tests','golden','trajectories.npz'))
for k in z.files:
    if k.endswith('/status'): print(k, z[k])
print('ff10 diff', np.abs(z['stack/state/10'][:,:7]-G['stack/state/10'][:,:7]).max(axis=0))
for f in (60,240,360):
    st=z['stack/state/%d'%f]; print(f, 'y', np.round(st[:,1],3), 'x', np.round(st[:,0],3), 'z', np.round(st[:,2],3), 'speed', np.round(np.sqrt((st[:,7:10]**2).sum(1)),3))
print('ref240 y', np.round(G['stack/state/240'][:,1],3))
st=z['stack70/state/120']; print('stack70 y', np.round(st[:,1],3))
for k in ("w256/state/60","w256/state/120","wall/state/90","coin/state/60","spheres/state/120"):
    st=z[k]; print(k, 'ymin', st[:,1].min(), 'maxspeed', np.sqrt((st[:,7:10]**2).sum(1)).max())
