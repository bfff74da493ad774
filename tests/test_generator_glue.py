"""The glue of the DM-GAN generator loss around the matching loss (losses.py:348-354):
``F.interpolate(fake, size=224)`` (nearest) as a gather/scatter kernel pair pinned to torch's own op, and the
CLS-drop + 4-D region view feeding the stale 6-argument ``words_loss`` call."""
import importlib

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import damsm_oracle as O

pkg = importlib.import_module("t2i_clip-gan_b200")


def test_generator_regions_is_a_view_of_the_encoder_output():
    hid = torch.randn(3, 50, 512)                       # (B, R+1, D) as the ViT produces it
    rf = hid.permute(0, 2, 1)                           # encode_image_verbose returns the permuted view (model.py:46-48)
    ref = rf[:, :, 1:].reshape(-1, 512, 7, 7)           # losses.py:350
    got = pkg.generator_regions(rf)
    assert got.shape == ref.shape and torch.equal(got, ref)
    assert got.untyped_storage().data_ptr() == hid.untyped_storage().data_ptr()      # no copy
    with pytest.raises(ValueError):
        pkg.generator_regions(torch.randn(2, 512, 12))   # 11 regions: not a square grid


@pytest.mark.gpu
@pytest.mark.parametrize("shape,size,dtype", [((2, 3, 256, 256), 224, torch.float32), ((1, 3, 64, 64), 224, torch.float32),
                                              ((3, 3, 256, 256), 224, torch.bfloat16), ((2, 1, 17, 31), (9, 40), torch.float32),
                                              ((1, 2, 299, 299), 224, torch.float16)])
def test_clip_resize_matches_torch_interpolate(shape, size, dtype):
    g = torch.Generator(device="cuda").manual_seed(sum(shape))
    x = torch.randn(shape, device="cuda", generator=g).to(dtype).requires_grad_(True)
    ref = F.interpolate(x, size=size)                   # the reference op itself (default mode 'nearest')
    got = pkg.clip_resize(x, size)
    assert got.dtype == dtype and torch.equal(got, ref)
    dy = torch.randn(ref.shape, device="cuda", generator=g).to(dtype)
    gref, = torch.autograd.grad(ref, x, dy)
    ggot, = torch.autograd.grad(got, x, dy)
    assert torch.allclose(ggot.float(), gref.float(), rtol=1e-6, atol=1e-6)


@pytest.mark.gpu
def test_generator_step_damsm_term_image_side_gradients():
    """BASELINE configs[2] (B=10, T=77, R=49): regions come out of the encoder as (B, D, 50) with the CLS token, the text
    side is detached (trainer.py:338,345); the stale 6-argument call scores them; gradients reach the image side only."""
    B, T, R = 10, 77, 49
    x = O.make_inputs(B, T, R, seed=2028, class_ids=True, n_classes=4)
    o = O.words_loss(x["words"], x["regions"], x["mask"], x["labels"], x["class_ids"], 4.0, 5.0, 10.0)
    hid = torch.zeros(B, R + 1, 512, device="cuda")
    hid[:, 1:] = torch.tensor(x["regions"], device="cuda")
    hid.requires_grad_(True)
    regions = pkg.generator_regions(hid.permute(0, 2, 1))
    words = torch.tensor(x["words"], device="cuda").permute(0, 2, 1)
    l0, l1, _ = pkg.words_loss(regions, words, torch.arange(B, device="cuda"), torch.tensor(x["cap_len"]), x["class_ids"], B)
    ((l0 + l1) * 5.0).backward()
    assert abs(l0.item() - o["loss0"]) <= 1e-5 and abs(l1.item() - o["loss1"]) <= 1e-5
    g = hid.grad.cpu().numpy()
    assert np.abs(g[:, 0]).max() == 0.0                 # the CLS row gets no gradient
    assert np.abs(g[:, 1:] / 5.0 - o["dregions"]).max() <= 1e-5 * np.abs(o["dregions"]).max()
