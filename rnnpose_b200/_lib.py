"""ctypes binding of libb200pose.so (include/b200pose.h).  No CPU fallback: a missing library is an
ImportError-grade failure at first use, and every compute call needs a CUDA device."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200pose.so")
_lib = None

_vp, _i, _f, _d, _sz = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_size_t

# name -> (restype, argtypes); mirrors include/b200pose.h one to one
SIGNATURES = {
    "b200pose_version": (_i, []),
    "b200pose_error_string": (C.c_char_p, [_i]),
    "b200pose_set_option": (_i, [C.c_char_p, _i]),
    "b200pose_get_option": (_i, [C.c_char_p, C.POINTER(_i)]),
    "b200pose_option_count": (_i, []),
    "b200pose_option_name": (C.c_char_p, [_i]),
    "b200pose_packed_weights_bytes": (_sz, []),
    "b200pose_pack_weights": (_i, [C.POINTER(_vp), _vp, _vp]),
    "b200pose_pyramid_floats": (_sz, [_i, _i, _i]),
    "b200pose_corr_pyramid": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "b200pose_corr_pyramid_tc_workspace_bytes": (_sz, [_i, _i, _i]),
    "b200pose_corr_pyramid_tc": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "b200pose_corr_lookup": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "b200pose_context_init": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "b200pose_flow_init": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "b200pose_update_workspace_bytes": (_sz, [_i, _i, _i]),
    "b200pose_update_block": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "b200pose_conv_layer_workspace_bytes": (_sz, [_i, _i, _i]),
    "b200pose_conv_layer_info": (_i, [_i] + [C.POINTER(_i)] * 5),
    "b200pose_conv_layer": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "b200pose_upsample_weight": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "b200pose_encoder_packed_weights_bytes": (_sz, []),
    "b200pose_encoder_pack_weights": (_i, [C.POINTER(_vp), _vp, _vp]),
    "b200pose_encoder_workspace_bytes": (_sz, [_i, _i, _i]),
    "b200pose_image_encoder": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "b200pose_zoom_crop_workspace_bytes": (_sz, [_i]),
    "b200pose_zoom_crop": (_i, [_vp] * 5 + [_i] * 7 + [_f, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200pose_pose_metrics_workspace_bytes": (_sz, [_i, _i]),
    "b200pose_pose_metrics": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    "b200pose_lm_workspace_bytes": (_sz, [_i, _i, _i]),
    "b200pose_lm_solve": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _i, _d, _d, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200pose_lm_backward_workspace_bytes": (_sz, [_i, _i, _i]),
    "b200pose_lm_backward": (_i, [_vp] * 6 + [_i, _i, _i, _f, _d, _d, _vp, _vp, _vp, _sz, _vp]),
    "b200pose_cholesky_solve": (_i, [_vp, _vp, _vp, _i, _vp]),
    "b200pose_se3_retract": (_i, [_vp, _vp, _i, _vp]),
    "b200pose_refine_workspace_bytes": (_sz, [_i, _i, _i]),
    "b200pose_refine_iters": (_i, [_vp] * 9 + [_f, _i, _i, _i, _i, _i, _i, _d, _d, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200pose_refine_host_scratch_bytes": (_sz, [_i, _i, _i, _i]),
    "b200pose_refine_iters_host": (_i, [_vp] * 9 + [_f, _i, _i, _i, _i, _i, _i, _d, _d, _i, _vp, _sz, _vp]),
    "b200pose_refine_host_staging_bytes": (_sz, [_i, _i, _i]),
    "b200pose_context_gather_texels": (_i, [_vp, _i, _i, _i, _vp, _i]),
    "b200pose_refine_iters_host2": (_i, [_vp] * 9 + [_f, _i, _i, _i, _i, _i, _i, _d, _d, _i, _vp, _sz, _vp, _sz, _i, _vp]),
    "b200pose_debug_set_conv_events": (_i, [_vp, _vp]),
    "b200pose_refine_launch_count": (_i, [_i, _i, _i, _i, _i]),
}


def build(verbose: bool = False) -> str:
    """Compile the library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j", str(min(8, os.cpu_count() or 1))]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libb200pose.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout[-2000:])
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `make -C rnnpose_b200/csrc` (or __graft_entry__.build()). "
                               "There is no CPU fallback for the refinement kernels.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().b200pose_error_string(rc).decode()
        raise RuntimeError(f"{what} failed with code {rc}: {msg}")
