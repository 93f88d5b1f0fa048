"""Data files the kernels consume.  ``weights/gru_update.pth`` is the reference's shipped RAFT update-block checkpoint
(reference weights/gru_update.pth, loaded at model/CFNet.py:71-74), byte-identical; it is data, not
code.  ``weights/img_fea_enc.pth`` (reference weights/img_fea_enc.pth, model/CFNet.py:34-37) feeds the BasicEncoder kernels."""
from __future__ import annotations

import os
from typing import Dict

import torch

WEIGHTS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "weights")


def load_update_weights() -> Dict[str, torch.Tensor]:
    """cf_net.update_block state dict with the ``update_block.`` prefix stripped (keys 'encoder.convc1.weight', ...)."""
    sd = torch.load(os.path.join(WEIGHTS_DIR, "gru_update.pth"), map_location="cpu")
    return {k[len("update_block."):]: v.float() for k, v in sd.items()}


def load_encoder_weights() -> Dict[str, torch.Tensor]:
    """ImageFeaEncoder state dict (keys 'fnet.conv1.weight', ...), reference weights/img_fea_enc.pth."""
    sd = torch.load(os.path.join(WEIGHTS_DIR, "img_fea_enc.pth"), map_location="cpu")
    return {k: v.float() for k, v in sd.items()}
