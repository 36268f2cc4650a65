"""Golden vectors of the output stage after netG (blend with the static drawing + tensor2im).

Run in the build container only (needs /root/reference):   python tests/golden/make_compose_golden.py

  * `tensor2im` is the UNMODIFIED reference function, imported from /root/reference/Module2/util/util.py:9-29
    (applied per frame, as `visualizer.save_images` does, Module2/util/visualizer.py:16-52).
  * The blend is inline code of `GeomCGTIFWTestModel.forward` (Module2/models/geomcgt_ifw_test_model.py:297-300); that
    class cannot be constructed here (hard-coded .cuda(), absent checkpoints), so its three lines are executed here
    with the same torch ops on the seeded inputs.
The script asserts that oracle/netg_oracle.py reproduces both bit for bit and stores the outputs in compose_<case>.npz.
"""
import importlib.util
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import netg_oracle as O  # noqa: E402

COMPOSE_CASES = {"compose_line": (2, 1, 4101), "compose_cartoon": (1, 3, 4103)}  # name: (B, output_nc, seed)


def reference_tensor2im():
    spec = importlib.util.spec_from_file_location("ref_util", "/root/reference/Module2/util/util.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.tensor2im


def main():
    tensor2im = reference_tensor2im()
    for name, (B, onc, seed) in COMPOSE_CASES.items():
        fake, mask, motion, stat = O.make_compose_inputs(B, onc, seed)
        # geomcgt_ifw_test_model.py:297-300
        mask1 = F.grid_sample(mask, motion, align_corners=True)
        blended = ((fake / 2 + 0.5) * mask1 + (stat / 2 + 0.5) * (1 - mask1)) * 2 - 1
        img = np.stack([tensor2im(blended[i:i + 1]) for i in range(B)])       # reference function, one frame at a time
        img_plain = np.stack([tensor2im(fake[i:i + 1]) for i in range(B)])   # conversion without blend
        assert torch.equal(O.blend_foreground(fake, mask, motion, stat), blended)
        assert np.array_equal(O.tensor2im_batch(blended), img) and np.array_equal(O.tensor2im_batch(fake), img_plain)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), blended=blended.numpy(), image=img, image_plain=img_plain)
        print(name, blended.shape, img.shape, float(blended.mean()), int(img.astype(np.int64).sum()))


if __name__ == "__main__":
    main()
