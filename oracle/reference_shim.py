"""
ORACLE -- test infrastructure only (never imported by detectorch_b200/).

Imports the UNMODIFIED reference (`/root/reference/lib`) on CPU under torch 2.x /
numpy 2.x so its own code can pin the oracle and generate golden vectors.  Only
usable where /root/reference exists (this container, not the GPU box).

The six compat items are the ones SURVEY.md 8c lists:
  1. np.float / np.int / np.bool aliases      (generate_anchors.py:63-64,72)
  2. collections.Mapping alias                (collate_custom.py:15)
  3. stub pycocotools / pycocotools.mask      (result_utils.py:22)
  4. a `cppcuda_cffi.roialign` module with the cffi calling convention
     (roi_align.py:20,56-83) over the reference's own compiled CPU loop
     (oracle/_ref/libroialign_ref.so)
  5. a `utils_cython` package holding the patched-and-built cython_nms
     (oracle/_ref/) ahead of lib/ on sys.path (boxes.py:53-55)
  6. FPN configs without a pkl need roi_feature_channels=1024 (caller's job)
"""
import collections
import collections.abc
import ctypes
import os
import sys
import types

import numpy as np

REF = os.environ.get("DETECTORCH_REFERENCE", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_installed = False


def available():
    return os.path.isdir(os.path.join(REF, "lib")) and os.path.exists(os.path.join(_HERE, "_ref", "libroialign_ref.so"))


def staged_available():
    """oracle/_ref holds the reference's pure-Python host modules (reflib.zip) and its notebooks' code cells, staged by build_ref.sh where
    the reference tree exists; unlike /root/reference these travel to the GPU box."""
    return all(os.path.exists(os.path.join(_HERE, "_ref", f)) for f in ("reflib.zip", "notebook_cells.json")) and \
        any(f.startswith("cython_nms") for f in os.listdir(os.path.join(_HERE, "_ref")))


def install_compat(stub_plotting=False):
    """Compat items 1, 2, 3 and 5 (no sys.path change): what ANY import of the reference's host modules needs under numpy 2 / python 3.12."""
    for name, typ in (("float", float), ("int", int), ("bool", bool)):
        if not hasattr(np, name):
            setattr(np, name, typ)
    if not hasattr(collections, "Mapping"):
        collections.Mapping = collections.abc.Mapping
        collections.Sequence = collections.abc.Sequence
    stubs = ["pycocotools", "pycocotools.mask", "pycocotools.coco", "pycocotools.cocoeval"]
    if stub_plotting:
        stubs += ["matplotlib", "matplotlib.pyplot", "skimage", "skimage.io"]
    for m in stubs:
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = types.ModuleType(m)
                if "." in m:
                    setattr(sys.modules[m.rsplit(".", 1)[0]], m.rsplit(".", 1)[1], sys.modules[m])
    for m, attr in (("pycocotools.coco", "COCO"), ("pycocotools.cocoeval", "COCOeval")):      # `from pycocotools.coco import COCO` (json_dataset.py:37)
        if not hasattr(sys.modules[m], attr):
            setattr(sys.modules[m], attr, type(attr, (), {}))
    if "utils_cython" not in sys.modules:
        uc = types.ModuleType("utils_cython")
        uc.__path__ = [os.path.join(_HERE, "_ref")]
        sys.modules["utils_cython"] = uc
        bb = types.ModuleType("utils_cython.cython_bbox")
        bb.bbox_overlaps = None
        sys.modules["utils_cython.cython_bbox"] = bb
        import importlib
        sys.modules["utils_cython.cython_nms"] = importlib.import_module("utils_cython.cython_nms")


def install():
    """Make `import model.detector`, `import utils.boxes` ... resolve to the reference."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree or oracle/_ref not available")
    for name, typ in (("float", float), ("int", int), ("bool", bool)):
        if not hasattr(np, name):
            setattr(np, name, typ)
    if not hasattr(collections, "Mapping"):
        collections.Mapping = collections.abc.Mapping
        collections.Sequence = collections.abc.Sequence
    for m in ("pycocotools", "pycocotools.mask", "pycocotools.coco", "pycocotools.cocoeval"):
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = types.ModuleType(m)

    # (4) cffi-convention RoIAlign over the reference's compiled loop
    so = ctypes.CDLL(os.path.join(_HERE, "_ref", "libroialign_ref.so"))
    pkg = types.ModuleType("cppcuda_cffi")
    pkg.__path__ = []
    mod = types.ModuleType("cppcuda_cffi.roialign")

    def roi_align_forward_cpu(inp, rois, out, ph, pw, scale, sr):
        inp, rois = inp.contiguous(), rois.contiguous()
        assert out.is_contiguous()
        so.roi_align_forward_loop(ctypes.c_int(out.numel()), ctypes.c_void_p(inp.data_ptr()),
                                  ctypes.c_void_p(rois.data_ptr()), ctypes.c_float(scale),
                                  ctypes.c_int(inp.size(1)), ctypes.c_int(inp.size(2)), ctypes.c_int(inp.size(3)),
                                  ctypes.c_int(ph), ctypes.c_int(pw), ctypes.c_int(sr),
                                  ctypes.c_int(rois.size(1)), ctypes.c_void_p(out.data_ptr()))
        return 1

    mod.roi_align_forward_cpu = roi_align_forward_cpu
    pkg.roialign = mod
    sys.modules["cppcuda_cffi"] = pkg
    sys.modules["cppcuda_cffi.roialign"] = mod

    # (5) utils_cython package from oracle/_ref (cython_bbox is not on the hot path: stub)
    uc = types.ModuleType("utils_cython")
    uc.__path__ = [os.path.join(_HERE, "_ref")]
    sys.modules["utils_cython"] = uc
    bb = types.ModuleType("utils_cython.cython_bbox")
    bb.bbox_overlaps = None
    sys.modules["utils_cython.cython_bbox"] = bb
    import importlib
    sys.modules["utils_cython.cython_nms"] = importlib.import_module("utils_cython.cython_nms")

    sys.path.insert(0, os.path.join(REF, "lib"))
    _installed = True
