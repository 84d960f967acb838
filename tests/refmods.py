"""Loaders for the checkers used by the tests: the reference's own extension modules built into
oracle/_ref/ (TEST INFRASTRUCTURE, see oracle/build_ref.sh) under private names, so they never
shadow the product's drop-in modules."""
import importlib.machinery
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
_cache = {}


def have_ref(name: str) -> bool:
    return os.path.exists(os.path.join(REF_DIR, name + ".so"))


def load_ref(name: str):
    """name in {pointnet2_cuda, iou3d_cuda, roipool3d_cuda}; returns the module or None."""
    if name in _cache:
        return _cache[name]
    path = os.path.join(REF_DIR, name + ".so")
    mod = None
    if os.path.exists(path):
        import torch  # noqa: F401  (the extension links against libtorch)
        loader = importlib.machinery.ExtensionFileLoader(name, path)
        spec = importlib.util.spec_from_loader(name, loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
    _cache[name] = mod
    return mod
