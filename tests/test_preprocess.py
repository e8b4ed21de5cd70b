"""Host-side request builder against vectors generated from the reference's own tiler
(oracle/gen_preprocess_goldens.py executes image_processing_perception_lm_fast.py's functions unmodified)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from grasp_any_region_b200 import preprocess as pp
from grasp_any_region_b200.config import gar_tiny

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "preprocess_goldens.json")))


def test_canvas_choice_matches_reference():
    assert len(GOLD["canvas"]) > 300
    for c in GOLD["canvas"]:
        got = pp.select_canvas(c["w"], c["h"], 448, c["max_num_tiles"])
        assert got == (c["ncw"], c["nch"]), (c, got)
        assert got[0] * got[1] <= c["max_num_tiles"]


@pytest.mark.parametrize("case", GOLD["pixels"], ids=lambda c: f"{c['w']}x{c['h']}")
def test_tiles_bit_identical_to_reference(case):
    g = torch.Generator().manual_seed(case["seed"])
    img = torch.randint(0, 256, (3, case["h"], case["w"]), generator=g, dtype=torch.uint8)
    px, (ncw, nch) = pp.tile_image(img, case["tile"], case["max_num_tiles"])
    assert (ncw, nch) == (case["ncw"], case["nch"]) and list(px.shape) == case["shape"]
    assert hashlib.sha256(np.ascontiguousarray(px.numpy()).tobytes()).hexdigest() == case["sha256"]


def test_id_map_round_trips_through_the_models_decode():
    """2 id / 255 - 1 in bf16, decoded as modeling_gar.py:315-319 does, gives back the painted ids; first writer wins;
    the thumbnail and the tiles see the same regions."""
    spec = gar_tiny()
    h, w = 300, 520
    m0 = np.zeros((h, w), bool); m0[40:160, 100:300] = True
    m1 = np.zeros((h, w), bool); m1[100:280, 250:500] = True       # overlaps m0: m0 keeps the overlap
    idmap = pp.paint_id_map([m0, m1], spec.prompt_numbers)
    assert idmap[120, 260] == 0 and idmap[200, 400] == 1 and idmap[0, 0] == spec.prompt_numbers
    ncw, nch = pp.select_canvas(w, h, 56, 8)
    mv = pp.tile_id_map(torch.from_numpy(idmap), ncw, nch, 56).to(torch.bfloat16)
    ids = torch.clamp(torch.round((mv + 1) / 2 * 255), 0, spec.prompt_numbers).long()
    assert ids.shape == (1 + ncw * nch, 3, 56, 56) and set(ids.unique().tolist()) == {0, 1, spec.prompt_numbers}
    assert torch.equal(ids[:, 0], ids[:, 1]) and torch.equal(ids[:, 0], ids[:, 2])
    # tiles re-assembled row-major reproduce the NEAREST-resized canvas
    canvas = ids[1:, 0].reshape(nch, ncw, 56, 56).permute(0, 2, 1, 3).reshape(nch * 56, ncw * 56)
    ys = np.minimum(((np.arange(nch * 56) + 0.5) * h / (nch * 56)).astype(int), h - 1)
    xs = np.minimum(((np.arange(ncw * 56) + 0.5) * w / (ncw * 56)).astype(int), w - 1)
    assert np.array_equal(canvas.numpy(), idmap[ys][:, xs])


def test_bbox_is_inclusive_max_over_size():
    m = np.zeros((100, 200), bool)
    m[10:20, 30:50] = True
    assert pp.bbox_of(m) == (30 / 200, 10 / 100, 49 / 200, 19 / 100)


def test_build_request_layout():
    spec = gar_tiny()
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(600, 900, 3), dtype=np.uint8)
    m0 = np.zeros((600, 900), bool); m0[50:300, 100:400] = True
    m1 = np.zeros((600, 900), bool); m1[350:590, 500:880] = True
    req = pp.build_request(spec, img, [m0, m1], text_ids=list(range(1000, 1030)), max_num_tiles=4, dtype=torch.float32)
    ncw, nch = req["aspect_ratios"][0].tolist()
    T = 1 + ncw * nch
    assert req["pixel_values"].shape == (T, 3, spec.vision.img_size, spec.vision.img_size)
    assert req["global_mask_values"].shape == req["pixel_values"].shape
    ids = req["input_ids"][0]
    assert int((ids == spec.image_token_id).sum()) == T * spec.tokens_per_tile
    for k in range(2):
        pos = (ids == spec.crop_tokens_ids[k]).nonzero().flatten()
        assert pos.numel() == spec.tokens_per_tile and int(pos.max() - pos.min()) + 1 == spec.tokens_per_tile
        assert str(spec.crop_tokens_ids[k]) in req["bboxes"][0]
    assert float(req["pixel_values"].min()) >= -1.0 - 1e-6 and float(req["pixel_values"].max()) <= 1.0 + 1e-6
