"""Degenerate regimes for the host build of clip2d.cuh against the reference Clipper (see clip_fuzz.py): collapsed polygons,
identical / integer-shifted copies, few and many rays, large coordinates.  CPU only; prints one line per regime."""
import sys, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import clip_fuzz as cf
def report(name,a,b):
    r,h,st=cf.run(a,b)
    bad=np.nonzero(r.view(np.int32)!=h.view(np.int32))[0]
    print(f"{name}: pairs={len(a)} nonzero_ref={np.count_nonzero(r)} mismatches={len(bad)} overflow={np.count_nonzero(st==2)} status_other={np.count_nonzero((st!=0)&(st!=2))}")
    for i in bad[:3]: print("   pair",i,"ref",r[i],"ours",h[i],"st",st[i], "a",a[i].tolist()[:6],"...")
# tiny polygons (all vertices collapse), radius 0.001..2
for rad in (0.001, 0.6, 1.0, 1.6, 2.5):
    a,b=cf.make_pairs(100000, 32, rad, 0.3, 1, span=3); report(f"radius {rad} R32",a,b)
# identical polygons
a,b=cf.make_pairs(50000, 32, 10, 0.3, 2); report("identical",a,a.copy())
# integer-shifted copies (many collinear overlapping edges)
rng=np.random.default_rng(3)
sh=rng.integers(-3,4,(50000,1,2)); report("shifted copy",a,a+sh)
# 128 rays, large radius, heavy noise
a,b=cf.make_pairs(20000,128,40,0.6,4); report("R128 r40 noise .6",a,b)
a,b=cf.make_pairs(20000,128,6,0.3,5); report("R128 r6 (dense integer collisions)",a,b)
a,b=cf.make_pairs(20000,96,200,0.9,6); report("R96 r200 noise .9",a,b)
# few rays
for R in (3,4,5,8):
    a,b=cf.make_pairs(50000,R,8,0.5,7+R); report(f"R{R}",a,b)
# huge coordinates
a,b=cf.make_pairs(50000,32,12,0.2,20); a+=30000; b+=30000; report("coords ~30000",a,b)
