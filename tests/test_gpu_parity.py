"""-m gpu parity tests: CUDA path (through the public API -> C ABI) vs the oracles."""
import pytest
import torch

import helpers as h

gg = h.gg
pytestmark = pytest.mark.gpu


def _upstream_grads(H, W, seed=1, depth_alpha=True):
    g = torch.Generator().manual_seed(seed)
    Gc = torch.randn(3, H, W, generator=g)
    if not depth_alpha:
        return (Gc, None, None)
    return (Gc, torch.randn(1, H, W, generator=g) * 0.3, torch.randn(1, H, W, generator=g))


def test_cfg1_parity_vs_c_oracle():
    """BASELINE.json configs[0]: 10k random Gaussians, 512x512, SH degree 3; fwd + bwd."""
    st = gg.scenes.random_cloud(10_000)
    cam = gg.scenes.cfg1_camera()
    S = h.settings_for(cam, st, device=torch.device("cuda:0"))
    grads = _upstream_grads(cam.image_height, cam.image_width)
    got = h.run_cuda(S, st, grads)
    ref = h.run_c_oracle(S, st, grads)
    assert int((got["radii"] != ref["radii"]).sum()) <= 2
    h.assert_images_close(got, ref)
    h.assert_grads_close(got["grads"], ref["grads"])


def test_cfg1_training_style_grads():
    """Only the colour feeds the loss (as in s2/s3): depth/alpha grads arrive as None."""
    st = gg.scenes.random_cloud(4_000)
    cam = gg.scenes.cfg1_camera(384, 272)     # not a multiple of 16 in height
    S = h.settings_for(cam, st, device=torch.device("cuda:0"))
    grads = _upstream_grads(cam.image_height, cam.image_width, depth_alpha=False)
    got = h.run_cuda(S, st, grads)
    ref = h.run_c_oracle(S, st, grads)
    h.assert_images_close(got, ref)
    h.assert_grads_close(got["grads"], ref["grads"])
