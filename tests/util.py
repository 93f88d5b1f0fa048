import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def load_update_weights():
    """Shipped RAFT update-block weights (reference weights/gru_update.pth, a data file consumed by the
    kernels) with the ``update_block.`` prefix stripped."""
    sd = torch.load(os.path.join(GOLDEN, "weights", "gru_update.pth"), map_location="cpu")
    return {k[len("update_block."):]: v.float() for k, v in sd.items()}
