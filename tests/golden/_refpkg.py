"""Import single modules of the reference package (/root/reference/stardist) in a container that lacks its third-party
dependencies (tensorflow / csbdeep / scikit-image): `stardist` and its sub-packages are registered as bare namespace modules
(their __init__ files are NOT executed), third-party imports resolve to permissive stubs, csbdeep.utils gets the three helpers
the pure-numpy code paths use, and stardist.lib.stardist2d / stardist3d are the reference's own compiled extensions from
oracle/_ref.  Only used by the make_*.py fixture generators (build container; the fixtures travel, this does not)."""
import importlib, os, sys, types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/stardist"


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"): raise AttributeError(name)
        return lambda *a, **k: (_ for _ in ()).throw(RuntimeError("stubbed third-party function %s.%s called" % (self.__name__, name)))


def setup():
    if ROOT not in sys.path: sys.path.insert(0, ROOT)
    from stardist_b200 import utils as U
    from oracle import ref_ext
    for name in ("skimage", "skimage.measure", "skimage.draw", "skimage.segmentation", "skimage.morphology", "skimage.transform",
                 "csbdeep", "csbdeep.utils", "csbdeep.utils.tf", "csbdeep.utils.six", "csbdeep.models", "csbdeep.internals", "tensorflow", "keras"):
        sys.modules.setdefault(name, _Stub(name))
    sys.modules["csbdeep.utils"].__path__ = []
    import pathlib
    sys.modules["csbdeep.utils.six"].Path = pathlib.Path
    cu = sys.modules["csbdeep.utils"]
    cu._raise = U._raise; cu.axes_check_and_normalize = U.axes_check_and_normalize; cu.axes_dict = U.axes_dict
    for pkg, sub in (("stardist", ""), ("stardist.geometry", "geometry"), ("stardist.lib", "lib"), ("stardist.models", "models")):
        m = types.ModuleType(pkg); m.__path__ = [os.path.join(REF, sub)] if sub else [REF]; m.__package__ = pkg
        sys.modules[pkg] = m
    sys.modules["stardist.lib.stardist2d"] = ref_ext.stardist2d()
    sys.modules["stardist.lib.stardist3d"] = ref_ext.stardist3d()
    # the geometry sub-package re-exports its two modules (stardist/geometry/__init__.py)
    g = sys.modules["stardist.geometry"]
    for sub in ("geom2d", "geom3d"):
        mod = importlib.import_module("stardist.geometry." + sub)
        for k in dir(mod):
            if not k.startswith("_"): setattr(g, k, getattr(mod, k))


def load(modname):
    """e.g. load('stardist.nms') -> the reference's module object"""
    setup()
    return importlib.import_module(modname)
