"""tests/golden/rays3d.npz: vertices / faces / volume / surface of the reference's ray classes (stardist/rays3d.py imported
standalone from /root/reference -- it only needs numpy and scipy).  Run in the build container: python tests/golden/make_rays3d.py"""
import importlib.util, json, os
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_rays3d", "/root/reference/stardist/rays3d.py")
ref = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref)

SPECS = ([("Rays_GoldenSpiral", dict(n=n, anisotropy=a)) for n in (4, 5, 14, 22, 32, 65, 70, 96, 100, 187) for a in (None, (2, 1, 1), (1, 1.5, 3))] +
         [("Rays_Cartesian", dict(n_rays_x=x, n_rays_z=z)) for x, z in ((11, 5), (8, 4), (16, 9), (5, 3))] +
         [("Rays_Tetra", dict(n_level=l)) for l in (1, 2, 3, 4)] + [("Rays_Octo", dict(n_level=l)) for l in (1, 2, 3, 4)])

if __name__ == "__main__":
    out = {"specs": np.frombuffer(json.dumps(SPECS).encode(), np.uint8)}
    rng = np.random.default_rng(0)
    for i, (name, kw) in enumerate(SPECS):
        r = getattr(ref, name)(**kw)
        d = rng.uniform(.5, 2., (3, len(r.vertices)))
        out["%d/vertices" % i] = r.vertices; out["%d/faces" % i] = r.faces.astype(np.int32)
        out["%d/dist" % i] = d; out["%d/volume" % i] = r.volume(d); out["%d/surface" % i] = r.surface(d)
        out["%d/weights" % i] = r.dist_loss_weights((2, 1, 1))
        scaled = r.copy(scale=(.5, 1, 2))
        out["%d/scaled" % i] = scaled.vertices
    np.savez_compressed(os.path.join(HERE, "rays3d.npz"), **out)
    print(len(SPECS), "ray sets,", os.path.getsize(os.path.join(HERE, "rays3d.npz")), "bytes")
