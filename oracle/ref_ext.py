"""Load the reference's compiled CPython extension modules from oracle/_ref by path
(importing the `stardist` package itself needs csbdeep/skimage/TF).  TEST INFRASTRUCTURE ONLY."""
import importlib.util, os, ctypes

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")


def _load(name):
    path = os.path.join(_REF, name + ".so")
    if not os.path.exists(path):
        raise FileNotFoundError("%s missing: run `make -C oracle ref` where /root/reference exists" % path)
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m

_cache = {}

def stardist2d():
    if "2d" not in _cache: _cache["2d"] = _load("stardist2d")
    return _cache["2d"]

def stardist3d():
    if "3d" not in _cache: _cache["3d"] = _load("stardist3d")
    return _cache["3d"]

def sdref():
    if "shim" not in _cache: _cache["shim"] = ctypes.CDLL(os.path.join(_REF, "libsdref.so"))
    return _cache["shim"]

def available():
    return all(os.path.exists(os.path.join(_REF, f)) for f in ("stardist2d.so", "stardist3d.so"))
