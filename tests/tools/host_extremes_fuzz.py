"""Host build of the 3-D NMS arithmetic and of the rendering rule (tests/hostcheck) against the reference extension at the
extremes: tiny polyhedra (distances 1e-3 ... 1.6), distances clamped at 1e-3, coordinates in the thousands, polyhedra
larger than the volume.  CPU only, OMP_NUM_THREADS=1; prints one line per case, all counts must be 0."""
import ctypes, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases
from oracle import ref_ext
hc=ctypes.CDLL(os.path.join(ROOT,"tests/hostcheck/_build/libhostcheck.so")); P=ctypes.c_void_p
hc.hc_nms3d_serial.argtypes = [P, P, P, P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, P, P]
hc.hc_polyhedron_to_label.argtypes=[P,P,P,P,ctypes.c_int,ctypes.c_int,ctypes.c_int,P]+[ctypes.c_int]*4+[P]
ext=ref_ext.stardist3d()
def nms(name, rays, pts, d, nthr):
    v=np.ascontiguousarray(rays.vertices,np.float32); f=np.ascontiguousarray(rays.faces,np.int32); n=len(d); R=len(v)
    s=np.ascontiguousarray(np.sort(np.random.default_rng(0).uniform(.5,1,n))[::-1],np.float32)
    want=ext.c_non_max_suppression_inds(d,pts,v,f,s,1,1,0,np.float32(nthr))
    keep=np.zeros(n,np.uint8); sc=np.zeros(5,np.int32)
    hc.hc_nms3d_serial(d.ctypes.data,pts.ctypes.data,v.ctypes.data,f.ctypes.data,n,R,len(f),ctypes.c_float(nthr),1,1,0,keep.ctypes.data,sc.ctypes.data)
    print("NMS",name,"n",n,"kept",int(want.sum()),"mismatches",int((keep.astype(bool)!=want).sum()),"stages",[int(x) for x in sc],flush=True)
def paint(name, rays, pts, d, shape):
    v=np.ascontiguousarray(rays.vertices,np.float32); f=np.ascontiguousarray(rays.faces,np.int32); n=len(d); R=len(v)
    labels=np.arange(1,n+1,dtype=np.int32); res=[]
    for mode in (0,1,2,3):
        want=ext.c_polyhedron_to_label(d,pts,v,f,labels,mode,0,0,0,shape)
        got=np.zeros(shape,np.int32)
        hc.hc_polyhedron_to_label(d.ctypes.data,pts.ctypes.data,v.ctypes.data,f.ctypes.data,n,R,len(f),labels.ctypes.data,*shape,mode,got.ctypes.data)
        res.append((int((want!=got).sum()),int((want>0).sum())))
    print("PAINT",name,res,flush=True)
rng=np.random.default_rng(1)
rays=cases.rays_golden_spiral(32,None); rays96=cases.rays_golden_spiral(96,(2,1,1))
# tiny polyhedra
n=400
pts=np.ascontiguousarray(np.stack([rng.integers(2,14,n),rng.integers(2,16,n),rng.integers(2,18,n)],1),np.float32)
for lo,hi in ((0.001,0.01),(0.2,0.8),(0.5,1.6)):
    d=np.ascontiguousarray(rng.uniform(lo,hi,(n,1))*(1+0.3*rng.uniform(-1,1,(n,32))),np.float32)
    nms(f"tiny {lo}-{hi}",rays,pts,d,0.3); paint(f"tiny {lo}-{hi}",rays,pts[:40],d[:40],(16,18,20))
# clamp 1e-3 exactly on some rays
d=np.ascontiguousarray(np.maximum(1e-3, rng.uniform(-1,2,(n,32))),np.float32)
nms("clamped 1e-3",rays,pts,d,0.3); paint("clamped 1e-3",rays,pts[:40],d[:40],(16,18,20))
# large coordinates
pts2=np.ascontiguousarray(np.stack([rng.integers(1000,1016,n),rng.integers(2000,2020,n),rng.integers(3000,3024,n)],1),np.float32)
d=np.ascontiguousarray(rng.uniform(2,5,(n,1))*(1+0.3*rng.uniform(-1,1,(n,96))),np.float32)
nms("coords 1000-3000 r96 aniso",rays96,pts2,d,0.3)
pts3=np.ascontiguousarray(np.stack([rng.uniform(230,246,30),rng.uniform(240,260,30),rng.uniform(220,240,30)],1),np.float32)
paint("coords ~240 float centres",rays96,pts3,d[:30],(256,270,250))
pts4=np.round(pts3)
paint("coords ~240 int centres",rays96,np.ascontiguousarray(pts4,np.float32),d[:30],(256,270,250))
# huge polyhedra relative to volume (clipping on all sides)
pts5=np.ascontiguousarray(np.stack([rng.uniform(0,20,10),rng.uniform(0,24,10),rng.uniform(0,28,10)],1),np.float32)
d5=np.ascontiguousarray(rng.uniform(15,30,(10,1))*(1+0.3*rng.uniform(-1,1,(10,32))),np.float32)
paint("huge clipped",rays,pts5,d5,(20,24,28))
# (centres outside the volume are left out: for some such inputs the reference's c_polyhedron_to_label itself does not return)
