"""Runs the UNMODIFIED reference (/root/reference) on CPU in the build container.

Only usable where /root/reference exists (never on the GPU box, never imported by tests):
``make_golden.py`` uses it to produce the committed fixtures.  Follows SURVEY.md Appendix B:
three in-memory stub modules (yacs, easydict, transforms3d -- not installed here), the reference
tree on sys.path, and an injected analytic renderer.
"""
import os
import sys
import types

REF = "/root/reference"


def install_stubs():
    if "yacs" not in sys.modules:
        class CfgNode(dict):
            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError:
                    raise AttributeError(k)

            def __setattr__(self, k, v):
                self[k] = v
        yacs = types.ModuleType("yacs"); cfgm = types.ModuleType("yacs.config")
        cfgm.CfgNode = CfgNode; yacs.config = cfgm
        sys.modules["yacs"] = yacs; sys.modules["yacs.config"] = cfgm
    if "easydict" not in sys.modules:
        class EasyDict(dict):
            def __init__(self, d=None, **kw):
                super().__init__()
                d = dict(d or {}); d.update(kw)
                for k, v in d.items():
                    self[k] = v

            def __setitem__(self, k, v):
                if isinstance(v, dict) and not isinstance(v, EasyDict):
                    v = EasyDict(v)
                super().__setitem__(k, v)

            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError:
                    raise AttributeError(k)

            def __setattr__(self, k, v):
                self[k] = v
        ed = types.ModuleType("easydict"); ed.EasyDict = EasyDict
        sys.modules["easydict"] = ed
    if "transforms3d" not in sys.modules:
        t3 = types.ModuleType("transforms3d")
        t3.quaternions = types.ModuleType("transforms3d.quaternions")
        t3.euler = types.ModuleType("transforms3d.euler")
        sys.modules["transforms3d"] = t3
        sys.modules["transforms3d.quaternions"] = t3.quaternions
        sys.modules["transforms3d.euler"] = t3.euler
    sys.dont_write_bytecode = True
    for p in (REF, os.path.join(REF, "thirdparty")):
        if p not in sys.path:
            sys.path.insert(0, p)


def motion_cfg(iter_count=4, optim_iter_count=3, render_iter_count=1):
    from easydict import EasyDict
    return EasyDict(IS_CALIBRATED=True, RESCALE_IMAGES=False, ITER_COUNT=iter_count,
                    RENDER_ITER_COUNT=render_iter_count, TRAIN_FLOW_WEIGHT=0.5, TRAIN_REPROJ_WEIGHT=0,
                    OPTIM_ITER_COUNT=optim_iter_count, FLOW_NET="raft", ONLINE_CROP=True,
                    raft=dict(small=False, fea_net="default", mixed_precision=True,
                              pretrained_model="x", input_dim=3, iters=1))


def build_reference_refiner(renderer, H, W, **cfgkw):
    """reference model.PoseRefiner.PoseRefiner on CPU with shipped weights."""
    install_stubs()
    import torch
    from config.default import get_cfg
    from model.PoseRefiner import PoseRefiner
    get_cfg().merge({"zoom_crop_size": [H, W], "render_image_size": [H, W]}, "BASIC")
    net = PoseRefiner(motion_cfg(**cfgkw), renderer=renderer)
    net.eval()
    return net


def install_eval_stubs():
    """Extra in-memory stubs so that the reference's evaluator (utils/eval_metric.py) imports unmodified:
    plyfile / open3d / matplotlib (not installed; unused by the metric methods), a namespace `data` package (its
    __init__ pulls the whole dataset stack), and thirdparty.nn._ext -- the compiled CUDA extension -- whose one entry
    point is restated in numpy from thirdparty/nn/src/nearest_neighborhood.cu:48-80 (float32 squared distances summed
    x, y, z; strict '<' so the FIRST minimum wins).  thirdparty/nn/nn_utils.py itself runs unmodified on top of it."""
    install_stubs()
    import ctypes
    import numpy as np

    def mod(name, **attrs):
        if name in sys.modules:
            m = sys.modules[name]
        else:
            m = types.ModuleType(name)
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    mod("plyfile", PlyData=object)
    mod("open3d")
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl = mod("matplotlib")
        mpl.cm = mod("matplotlib.cm", get_cmap=lambda *a, **k: None)
        mpl.pyplot = mod("matplotlib.pyplot")
        mpl.patches = mod("matplotlib.patches")
    try:
        import scipy.misc  # noqa: F401
    except ImportError:
        import scipy
        scipy.misc = mod("scipy.misc")
    q = sys.modules["transforms3d.quaternions"]
    for n in ("mat2quat", "quat2mat", "qmult"):
        if not hasattr(q, n):
            setattr(q, n, None)
    if "data" not in sys.modules:
        d = mod("data")
        d.__path__ = [os.path.join(REF, "data")]

    class _FFI:
        @staticmethod
        def cast(ctype, addr):
            return (ctype, int(addr))

    class _Lib:
        @staticmethod
        def findNearestPointIdxLauncher(ref_ptr, que_ptr, idx_ptr, b, pn1, pn2, dim, exclude_self):
            assert b == 1 and not exclude_self
            ref = np.ctypeslib.as_array(ctypes.cast(ref_ptr[1], ctypes.POINTER(ctypes.c_float)), shape=(pn1, dim))
            que = np.ctypeslib.as_array(ctypes.cast(que_ptr[1], ctypes.POINTER(ctypes.c_float)), shape=(pn2, dim))
            idx = np.ctypeslib.as_array(ctypes.cast(idx_ptr[1], ctypes.POINTER(ctypes.c_int32)), shape=(pn2,))
            for i in range(pn2):
                d = np.zeros(pn1, np.float32)
                for k in range(dim):                       # (x1-x2)^2 + (y1-y2)^2 + (z1-z2)^2 in float32, in this order
                    diff = ref[:, k] - que[i, k]
                    d = d + diff * diff
                idx[i] = int(np.argmin(d))                  # first minimum, as the strict '<' scan

    mod("thirdparty.nn._ext", lib=_Lib, ffi=_FFI)


def reference_evaluator(model_pts, diameter_m):
    """utils.eval_metric.LineMODEvaluator with its metric methods untouched; the constructor (which reads a .ply from
    EXPDATA, absent here) is bypassed and the attributes it would set are filled from the arguments."""
    install_eval_stubs()
    import numpy as np
    import utils.eval_metric as EM
    ev = EM.LineMODEvaluator.__new__(EM.LineMODEvaluator)
    ev.class_name = "synthetic"
    ev.model = np.asarray(model_pts, np.float32)
    ev.diameter = float(diameter_m)
    for name in ("proj2d", "add", "adds", "add2", "add5", "cmd5", "icp_proj2d", "icp_add", "icp_cmd5", "mask_ap", "pose_preds"):
        setattr(ev, name, [])
    return ev, EM
