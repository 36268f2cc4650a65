"""Output stage after netG (blend + tensor2im): oracle against the golden vectors made from the reference (CPU),
CUDA kernel against the oracle and the golden vectors through the C ABI (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import netg_oracle as O
from tests.golden.make_compose_golden import COMPOSE_CASES


@pytest.mark.parametrize("name", list(COMPOSE_CASES))
def test_oracle_reproduces_reference_compose_golden(name, golden_dir):
    B, onc, seed = COMPOSE_CASES[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    fake, mask, motion, stat = O.make_compose_inputs(B, onc, seed)
    blended = O.blend_foreground(fake, mask, motion, stat)
    assert np.array_equal(blended.numpy(), g["blended"])           # same torch ops: bit-exact
    assert np.array_equal(O.tensor2im_batch(blended), g["image"])
    assert np.array_equal(O.tensor2im_batch(fake), g["image_plain"])
    assert g["image"].shape == (B, 256, 256, 3) and g["image"].dtype == np.uint8


def test_compose_wrapper_rejects_cpu_and_partial_arguments():
    from animateportrait_b200.compose import blend_and_convert
    with pytest.raises(RuntimeError, match="CUDA"):
        blend_and_convert(torch.zeros(1, 1, 256, 256))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(COMPOSE_CASES))
def test_compose_kernel_matches_golden(name, golden_dir):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from animateportrait_b200.compose import blend_and_convert
    dev = torch.device("cuda", 0)
    B, onc, seed = COMPOSE_CASES[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    fake, mask, motion, stat = (t.to(dev) for t in O.make_compose_inputs(B, onc, seed))
    blended, image = blend_and_convert(fake, mask, motion, stat)
    torch.cuda.synchronize()
    # elementwise fp32 in the reference's op order: bit-exact up to the last ulp of the bilinear mask sample
    assert np.abs(blended.cpu().numpy() - g["blended"]).max() <= 2e-6
    d = np.abs(image.cpu().numpy().astype(np.int16) - g["image"].astype(np.int16))
    assert d.max() <= 1 and (d > 0).mean() <= 1e-3   # a truncation boundary may flip on a last-ulp difference
    # conversion only (tensor2im of the raw frame): no interpolation involved, must be exact
    _, plain = blend_and_convert(fake, want_blended=False)
    assert np.array_equal(plain.cpu().numpy(), g["image_plain"])
    with pytest.raises(RuntimeError, match="together"):
        blend_and_convert(fake, mask=mask)
