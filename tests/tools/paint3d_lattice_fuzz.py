"""Quantify the 3-D rendering caveat (DESIGN.md 5, 3D labels): k_paint3d rule (host build, tests/hostcheck) vs the reference c_polyhedron_to_label
on polyhedra with integer centres and integer / half-integer / fractional ray lengths.  CPU only (needs oracle/_ref)."""
import ctypes, sys, numpy as np
ROOT='/root/repo'; sys.path.insert(0, ROOT); sys.path.insert(0, ROOT+'/tests/golden')
import cases
from oracle import pipeline3d
hc = ctypes.CDLL(ROOT+"/tests/hostcheck/_build/libhostcheck.so")
hc.hc_paint3d.argtypes = [ctypes.c_void_p]*4 + [ctypes.c_int]*5 + [ctypes.c_void_p]
def paint(dist, point, rays, shape):
    v=np.ascontiguousarray(rays.vertices,np.float32); f=np.ascontiguousarray(rays.faces,np.int32)
    d=np.ascontiguousarray(dist,np.float32); c=np.ascontiguousarray(point,np.float32); out=np.zeros(shape,np.uint8)
    hc.hc_paint3d(d.ctypes.data,c.ctypes.data,v.ctypes.data,f.ctypes.data,len(v),len(f),shape[0],shape[1],shape[2],out.ctypes.data); return out
rng=np.random.default_rng(1)
shape=(40,44,48)
for label, gen in (("integer centre, integer dist (all rays equal)", lambda n: np.full(n, rng.integers(4,14)).astype(np.float32)),
                   ("integer centre, integer dists (random per ray)", lambda n: rng.integers(4,14,n).astype(np.float32)),
                   ("integer centre, half-integer dists", lambda n: (rng.integers(8,28,n)/2).astype(np.float32)),
                   ("integer centre, fractional dists", lambda n: rng.uniform(4,14,n).astype(np.float32))):
    worst=0; tot_diff=0; tot_vox=0; npoly=0; extra=0; missing=0
    for n_rays in (32, 64, 96):
        for aniso in (None, (2,1,1)):
            rays=cases.rays_golden_spiral(n_rays, aniso)
            for _ in range(12):
                dist=gen(n_rays); point=np.array([rng.integers(14,26),rng.integers(14,30),rng.integers(14,34)])
                ours=paint(dist,point,rays,shape)
                ref=pipeline3d.polyhedron_to_label(dist[None],point[None],rays,shape,np.ones(1))
                nd=int(np.count_nonzero((ours>0)!=(ref>0))); nv=int(np.count_nonzero(ref))
                extra+=int(np.count_nonzero((ours>0)&(ref==0))); missing+=int(np.count_nonzero((ours==0)&(ref>0)))
                worst=max(worst, nd/max(nv,1)); tot_diff+=nd; tot_vox+=nv; npoly+=1
    print("%-48s polyhedra %3d  voxels %8d  differing %4d (ours only %d, reference only %d)  worst per polyhedron %.2e"%(label,npoly,tot_vox,tot_diff,extra,missing,worst))
