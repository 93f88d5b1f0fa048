import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def load_update_weights():
    """Shipped RAFT update-block weights (rnnpose_b200/weights/gru_update.pth = reference weights/gru_update.pth, a data
    file consumed by the kernels) with the ``update_block.`` prefix stripped."""
    from rnnpose_b200.assets import load_update_weights as _l
    return _l()
