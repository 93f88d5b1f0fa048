"""Mirror of the reference's ``model.CFNet.ImageFeaEncoder`` (reference model/CFNet.py:26-49) on the library's kernels:
same state-dict keys (``fnet.conv1.weight`` ... ``fnet.conv2.bias``, 32 tensors; ``weights/img_fea_enc.pth`` loads with
``strict=True``), same ``forward(image1, image2) -> (fmap1, fmap2)``.  Inference only.  The nn.Conv2d modules only hold the
parameters (thirdparty/raft/extractor.py:118-232 layout); the arithmetic is ``ops.image_encoder`` (csrc/encoder.cu)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import assets, ops


def _block(cin, cout, stride):
    b = nn.Module()
    b.conv1 = nn.Conv2d(cin, cout, 3, padding=1, stride=stride)
    b.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
    if stride != 1:
        b.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride))     # key 'downsample.0'; norm3 has no parameters
    return b


class _BasicEncoderParams(nn.Module):
    def __init__(self, output_dim=256, input_dim=3):
        super().__init__()
        self.conv1 = nn.Conv2d(input_dim, 64, 7, stride=2, padding=3)
        self.layer1 = nn.Sequential(_block(64, 64, 1), _block(64, 64, 1))
        self.layer2 = nn.Sequential(_block(64, 96, 2), _block(96, 96, 1))
        self.layer3 = nn.Sequential(_block(96, 128, 2), _block(128, 128, 1))
        self.conv2 = nn.Conv2d(128, output_dim, 1)


class ImageFeaEncoder(nn.Module):
    def __init__(self, input_dim=3, output_dim=256, load_shipped_weights=True):
        super().__init__()
        if input_dim != 3 or output_dim != 256:
            raise NotImplementedError("the kernels implement the shipped configuration (3 -> 256)")
        self.fnet = _BasicEncoderParams(output_dim, input_dim)
        if load_shipped_weights:                       # the reference loads weights/img_fea_enc.pth in its constructor (:33-37)
            self.load_state_dict(assets.load_encoder_weights(), strict=True)
        self._packed = None
        self._key = None
        self._ws = None

    def packed_weights(self) -> torch.Tensor:
        sd = self.state_dict()
        key = tuple((k, v.data_ptr(), v._version, str(v.device)) for k, v in sd.items())
        if self._packed is None or key != self._key:
            self._packed = ops.encoder_pack_weights(dict(sd), self.fnet.conv1.weight.device)
            self._key = key
        return self._packed

    @torch.no_grad()
    def forward(self, image1, image2):
        return ops.image_encoder(self.packed_weights(), image1.float().contiguous(), image2.float().contiguous())
