"""Import single modules of the reference package (/root/reference/stardist) in a container that lacks its third-party
dependencies (tensorflow / csbdeep / scikit-image): `stardist` and its sub-packages are registered as bare namespace modules
(their __init__ files are NOT executed), third-party imports resolve to permissive stubs, csbdeep.utils gets the three helpers
the pure-numpy code paths use, and stardist.lib.stardist2d / stardist3d are the reference's own compiled extensions from
oracle/_ref.  Only used by the make_*.py fixture generators (build container; the fixtures travel, this does not)."""
import importlib, os, sys, types
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/stardist"


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"): raise AttributeError(name)
        return lambda *a, **k: (_ for _ in ()).throw(RuntimeError("stubbed third-party function %s.%s called" % (self.__name__, name)))


class _Region:
    def __init__(self, label, sl, image):
        self.label = label; self.slice = sl; self.image = image
        self.bbox = tuple(s.start for s in sl) + tuple(s.stop for s in sl)


def _regionprops(label_image):
    """stand-in for skimage.measure.regionprops restricted to what stardist/big.py reads (label, bbox, image): one region
    per label present, ascending label order, bbox = (min_0, min_1[, min_2], max_0, max_1[, max_2]) half open"""
    from scipy.ndimage import find_objects
    label_image = np.asarray(label_image)
    return [_Region(i + 1, sl, label_image[sl] == i + 1) for i, sl in enumerate(find_objects(label_image)) if sl is not None]


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
    sys.modules["skimage.measure"].regionprops = _regionprops
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


class _Any:
    """absorbs any construction / call / attribute access (keras layers, tensorflow symbols at import time)"""
    def __init__(self, *a, **k): pass
    def __call__(self, *a, **k): return _Any()
    def __getattr__(self, n): return _Any()


class _BaseConfig:
    """the few attributes csbdeep.models.BaseConfig.__init__ provides that the reference's Config2D / Config3D read back"""
    def __init__(self, axes='YX', n_channel_in=1, n_channel_out=1, allow_new_parameters=False, **kwargs):
        ax = ''.join(a for a in str(axes).upper() if a != 'C')
        self.n_dim = len(ax); self.axes = ax + 'C'
        self.n_channel_in = int(max(1, n_channel_in)); self.n_channel_out = int(max(1, n_channel_out))
        self.train_checkpoint = 'weights_best.h5'; self.train_checkpoint_last = 'weights_last.h5'; self.train_checkpoint_epoch = 'weights_now.h5'
        for k, v in kwargs.items(): setattr(self, k, v)

    def update_parameters(self, allow_new=False, **kwargs):
        for k, v in kwargs.items(): setattr(self, k, v)

    def is_valid(self, return_invalid=False):
        return (True, ()) if return_invalid else True


def load_models():
    """-> (stardist.models.model2d, stardist.models.model3d) of the reference; only their pure host logic is usable
    (e.g. StarDist2D._instances_from_prediction called on a stand-in `self` with .config.grid and .thresholds)"""
    setup()
    for name in ("csbdeep.models.base_model", "csbdeep.internals.predict", "csbdeep.internals.train", "csbdeep.data", "csbdeep.internals",
                 "tensorflow", "csbdeep.internals.blocks", "csbdeep.internals.nets", "csbdeep.models", "csbdeep.utils.tf", "csbdeep.internals.probability"):
        sys.modules[name] = _Stub(name)
    sys.modules["csbdeep.models.base_model"].BaseModel = type("BaseModel", (object,), {})
    sys.modules["csbdeep.models"].BaseConfig = _BaseConfig
    sys.modules["csbdeep.data"].Resizer = type("Resizer", (object,), {})
    sys.modules["csbdeep.internals.train"].RollingSequence = type("RollingSequence", (object,), {})
    tfm = sys.modules["csbdeep.utils.tf"]

    def keras_import(sub=None, *names):
        if not names:
            k = _Any(); k.__dict__['__version__'] = '2.15.0'
            return k
        cls = [type(n, (object,), {}) for n in names]
        return cls[0] if len(cls) == 1 else tuple(cls)
    tfm.keras_import = keras_import; tfm.IS_TF_1 = False; tfm.IS_KERAS_3_PLUS = False; tfm.BACKEND = _Any(); tfm.export_SavedModel = _Any()
    tfm.CARETensorBoard = type("CARETensorBoard", (object,), {}); tfm.CARETensorBoardImage = type("CARETensorBoardImage", (object,), {})
    cu = sys.modules["csbdeep.utils"]; cu.backend_channels_last = lambda: True; cu.load_json = None; cu.save_json = None; cu.normalize = None
    return importlib.import_module("stardist.models.model2d"), importlib.import_module("stardist.models.model3d")
