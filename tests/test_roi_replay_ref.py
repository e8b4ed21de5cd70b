"""Pins oracle/roi_align_ref.py bit-exactly to the installed torchvision CPU operator, on the
call pattern the reference uses (modeling_gar.py:377-396: box pre-scaled, spatial_scale applied
again, 16x16 output, sampling_ratio 2, aligned)."""
import numpy as np
import pytest
import torch
import torchvision

from oracle.gar_oracle import roi_box_f32
from oracle.roi_align_ref import roi_align_ref

BOXES = [
    (0.0, 0.0, 1.0, 1.0),
    (0.1, 0.2, 0.7, 0.9),
    (0.3337, 0.0127, 0.3340, 0.99),       # degenerate thin box
    (0.9990234375, 0.9990234375, 0.9990234375, 0.9990234375),  # single pixel at the far corner
    (0.0, 0.5, 0.0625, 0.5625),
]


@pytest.mark.parametrize("grid", [(2, 2), (4, 4), (3, 2)])
@pytest.mark.parametrize("box", BOXES)
def test_numpy_roi_align_bit_exact(grid, box):
    ncw, nch = grid
    C = 24
    g = torch.Generator().manual_seed(7)
    fmap = torch.randn(1, C, nch * 16, ncw * 16, generator=g).bfloat16().float()
    roi, scale = roi_box_f32(box, nch * 16, ncw * 16)
    want = torchvision.ops.roi_align(fmap, roi[None], (16, 16), spatial_scale=scale, sampling_ratio=2,
                                     aligned=True)[0].numpy()
    got = roi_align_ref(fmap[0].numpy(), roi.numpy(), scale)
    assert got.dtype == np.float32
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_random_boxes_bit_exact():
    rng = np.random.default_rng(0)
    C = 8
    fmap = torch.randn(1, C, 64, 64).bfloat16().float()
    for _ in range(25):
        x1, y1 = rng.random(2) * 0.8
        box = (x1, y1, x1 + rng.random() * 0.2, y1 + rng.random() * 0.2)
        roi, scale = roi_box_f32(box, 64, 64)
        want = torchvision.ops.roi_align(fmap, roi[None], (16, 16), spatial_scale=scale, sampling_ratio=2,
                                         aligned=True)[0].numpy()
        got = roi_align_ref(fmap[0].numpy(), roi.numpy(), scale)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
