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
