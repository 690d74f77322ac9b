"""Exact (rational arithmetic) characterisation of the voxels on which render mode "full" may differ from the reference.

The reference labels a voxel iff  kernel || (hull && polyhedron)  (stardist3d_impl.cpp:1469-1476) with `hull` evaluated
on Qhull's facet planes in double.  polyhedron => hull holds in exact arithmetic, so the hull conjunct can only change
the result where Qhull's rounded plane puts a voxel that lies EXACTLY on a hull facet on the outside (|dist| ~ 1e-15;
its sign depends on Qhull's internal vertex numbering).  The product has no hull conjunct.  These helpers decide, with
fractions.Fraction on the float32 vertices the C code builds (centre + d * ray, float arithmetic), whether a voxel lies
exactly on the boundary of the convex hull -- the only place a difference is legitimate; the tests assert that every
differing voxel is such a voxel, labelled by the product and left out by the reference."""
from fractions import Fraction
import numpy as np
from scipy.spatial import ConvexHull


def polyhedron_vertices_f32(dist_row, center, rays_vertices):
    """stardist3d_impl.cpp polyhedron_polyverts: center + dist * vertex in float32"""
    d = np.asarray(dist_row, np.float32)[:, None]
    return (np.asarray(center, np.float32)[None] + d * np.asarray(rays_vertices, np.float32)).astype(np.float32)


def _orient(a, b, c, p):
    ax, ay, az = (b[i] - a[i] for i in range(3)); bx, by, bz = (c[i] - a[i] for i in range(3)); cx, cy, cz = (p[i] - a[i] for i in range(3))
    return ax * (by * cz - bz * cy) - ay * (bx * cz - bz * cx) + az * (bx * cy - by * cx)


def on_hull_boundary_exact(voxel, verts_f32):
    """True iff the integer voxel lies in the closed convex hull of the vertices and on at least one facet plane, exactly"""
    V = [[Fraction(float(x)) for x in v] for v in verts_f32]
    p = [Fraction(int(x)) for x in voxel]
    hull = ConvexHull(np.asarray(verts_f32, np.float64))
    inner = [sum(V[i][k] for i in hull.vertices) / len(hull.vertices) for k in range(3)]
    on_plane = False
    for tri in hull.simplices:
        a, b, c = (V[i] for i in tri)
        s_in = _orient(a, b, c, inner)
        s_p = _orient(a, b, c, p)
        if s_in == 0:
            continue                      # degenerate sliver reported by qhull's triangulation
        if s_p == 0:
            on_plane = True
        elif (s_p > 0) != (s_in > 0):
            return False                  # strictly outside this facet
    return on_plane


def assert_only_exact_hull_boundary_voxels_differ(got, want, dist, points, rays_vertices, max_voxels=64):
    """every voxel where the product's label map differs from the reference's is labelled by the product, unlabelled (or
    labelled as a single cover where the product sees an overlap) by the reference, and lies exactly on the hull boundary
    of a polyhedron that covers it"""
    idx = np.argwhere(got != want)
    assert len(idx) <= max_voxels, "%d voxels differ" % len(idx)
    for v in idx:
        g, w = int(got[tuple(v)]), int(want[tuple(v)])
        assert g != 0, ("voxel labelled by the reference only", tuple(v), g, w)
        ok = False
        for k in range(len(dist)):
            verts = polyhedron_vertices_f32(dist[k], points[k], rays_vertices)
            if np.any(np.rint(verts.min(0)) > v) or np.any(np.rint(verts.max(0)) < v):
                continue
            if on_hull_boundary_exact(v, verts):
                ok = True
                break
        assert ok, ("differing voxel is not on a hull facet", tuple(v), g, w)
    return len(idx)
