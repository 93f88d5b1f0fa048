"""Generates tests/golden/encoder.npz by EXECUTING the reference's ImageFeaEncoder (model/CFNet.py:26-49 with
thirdparty/raft/extractor.py and weights/img_fea_enc.pth, unmodified) on CPU fp32.  Run in the build container only:

    python tests/golden/make_golden_encoder.py

Cases: (a) two random 64x96 image pairs in [0, 255] (inputs stored); (b) the synthetic crop pair of scene 0 at 240x320 (inputs
regenerated bit-exactly from the seed by rnnpose_b200.synthetic; the reference feeds [0,1] images here);
(c) a 72x104 pair whose 1/2, 1/4 maps have odd sizes (36x52 -> 18x26 -> 9x13)."""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

import ref_harness as RH  # noqa: E402

RH.install_stubs()
from rnnpose_b200 import synthetic as S  # noqa: E402


def main():
    from model.CFNet import ImageFeaEncoder
    torch.set_num_threads(8)
    enc = ImageFeaEncoder().eval()
    out = {}
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        a = torch.rand(2, 3, 64, 96, generator=g) * 255; b = torch.rand(2, 3, 64, 96, generator=g) * 255
        f1, f2 = enc(a, b)
        out.update(a_img1=a.numpy(), a_img2=b.numpy(), a_f1=f1.float().numpy(), a_f2=f2.float().numpy())
        mb = S.make_batch([0], 240, 320, with_images=True)
        f1, f2 = enc(mb["syn_img"], mb["obs_img"])
        out.update(b_f1=f1.float().numpy(), b_f2=f2.float().numpy(), b_meta=np.array([0, 240, 320]))
        a = torch.rand(1, 3, 72, 104, generator=g) * 255; b = torch.rand(1, 3, 72, 104, generator=g) * 255
        f1, f2 = enc(a, b)
        out.update(c_img1=a.numpy(), c_img2=b.numpy(), c_f1=f1.float().numpy(), c_f2=f2.float().numpy())
    path = os.path.join(HERE, "encoder.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
