"""Ray factories for the 3D model (host side, computed once per model).

Mirrors stardist/rays3d.py: Rays_Base (:20-152), rays_from_json (:156-157, via a registry instead
of eval), Rays_Explicit (:162-168), Rays_GoldenSpiral (:337-373) and reorder_faces (:330-334).
Vertices are float32 (n,3) in (z,y,x) order, faces come from scipy.spatial.ConvexHull (the same
Qhull call the reference makes), re-oriented to point outward.
"""
import copy
import numpy as np
from scipy.spatial import ConvexHull


class Rays_Base(object):
    def __init__(self, **kwargs):
        self.kwargs = kwargs
        self._vertices, self._faces = self.setup_vertices_faces()
        self._vertices = np.asarray(self._vertices, np.float32)
        self._faces = np.asarray(self._faces, int)
        self._faces = np.asanyarray(self._faces)

    def setup_vertices_faces(self):
        raise NotImplementedError()

    @property
    def vertices(self):
        """read-only property"""
        return self._vertices.copy()

    @property
    def faces(self):
        """read-only property"""
        return self._faces.copy()

    def __getitem__(self, i):
        return self.vertices[i]

    def __len__(self):
        return len(self._vertices)

    def __repr__(self):
        def _conv(x):
            if isinstance(x, (tuple, list, np.ndarray)):
                return "_".join(_conv(_x) for _x in x)
            if isinstance(x, float):
                return "%.2f" % x
            return str(x)
        return "%s_%s" % (self.__class__.__name__, "_".join("%s_%s" % (k, _conv(v)) for k, v in sorted(self.kwargs.items())))

    def to_json(self):
        return {"name": self.__class__.__name__, "kwargs": self.kwargs}

    def copy(self, scale=(1, 1, 1)):
        """ returns a copy whose vertices are scaled by given factor"""
        scale = np.asarray(scale)
        assert scale.shape == (3,)
        res = copy.deepcopy(self)
        res._vertices *= scale[np.newaxis]
        return res


def reorder_faces(verts, faces):
    """reorder faces such that their orientation points outward"""
    def _single(face):
        return face[::-1] if np.linalg.det(verts[face]) > 0 else face
    return tuple(map(_single, faces))


class Rays_Explicit(Rays_Base):
    def __init__(self, vertices0, faces0):
        self.vertices0, self.faces0 = vertices0, faces0
        super().__init__(vertices0=list(vertices0), faces0=list(faces0))

    def setup_vertices_faces(self):
        return self.vertices0, self.faces0


class Rays_GoldenSpiral(Rays_Base):
    def __init__(self, n=70, anisotropy=None):
        if n < 4:
            raise ValueError("At least 4 points have to be given!")
        super().__init__(n=n, anisotropy=anisotropy if anisotropy is None else tuple(anisotropy))

    def setup_vertices_faces(self):
        n = self.kwargs["n"]
        anisotropy = self.kwargs["anisotropy"]
        if anisotropy is None:
            anisotropy = np.ones(3)
        else:
            anisotropy = np.array(anisotropy)
        # the smaller golden angle = 2pi * 0.3819...
        g = (3. - np.sqrt(5.)) * np.pi
        phi = g * np.arange(n)
        z = np.linspace(-1, 1, n)
        rho = np.sqrt(1. - z ** 2)
        verts = np.stack([z, rho * np.sin(phi), rho * np.cos(phi)]).T
        verts = verts / anisotropy
        hull = ConvexHull(verts)
        faces = reorder_faces(verts, hull.simplices)
        verts /= np.linalg.norm(verts, axis=-1, keepdims=True)
        return verts, faces


_RAY_CLASSES = {c.__name__: c for c in (Rays_Explicit, Rays_GoldenSpiral)}


def rays_from_json(d):
    name = d["name"]
    if name not in _RAY_CLASSES:
        raise ValueError("unknown / unsupported rays type '%s' (supported: %s)" % (name, sorted(_RAY_CLASSES)))
    return _RAY_CLASSES[name](**d["kwargs"])
