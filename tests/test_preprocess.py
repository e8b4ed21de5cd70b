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


# ---------------------------------------------------------------------------------------------------------------
# id map + boxes against the reference's own request builders (oracle/gen_idmap_goldens.py executes
# evaluation/eval_dataset.py `_parse_annotations` of both datasets unmodified)
# ---------------------------------------------------------------------------------------------------------------
IDGOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "idmap_goldens.json")))


def _rect_masks(seed, h, w, n):
    g = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        hh, ww = int(g.integers(h // 8, h // 2)), int(g.integers(w // 8, w // 2))
        y0, x0 = int(g.integers(0, h - hh)), int(g.integers(0, w - ww))
        m = np.zeros((h, w), dtype=bool)
        m[y0:y0 + hh, x0:x0 + ww] = True
        out.append(m)
    return out


@pytest.mark.parametrize("case", IDGOLD["single"], ids=lambda c: f"seed{c['seed']}")
def test_single_region_id_map_and_box_match_reference(case):
    spec = gar_tiny()
    m = _rect_masks(case["seed"], case["h"], case["w"], 1)
    idmap = pp.paint_id_map(m, spec.prompt_numbers, prompt_ids=[case["prompt"]])
    assert hashlib.sha256(np.ascontiguousarray(idmap).tobytes()).hexdigest() == case["sha256"]
    boxes = pp.region_boxes(spec, m, prompt_ids=[case["prompt"]])
    assert {k: list(map(float, v)) for k, v in boxes.items()} == case["bboxes"]


@pytest.mark.parametrize("case", IDGOLD["multi"], ids=lambda c: f"seed{c['seed']}")
def test_multi_region_quirks_match_reference(case):
    """the reference's multi-region builder gives EVERY region the last mask's box and paints in the
    visiting order of a set of strings; reference_quirks=True reproduces both, the default does neither."""
    spec = gar_tiny()
    masks = _rect_masks(case["seed"], case["h"], case["w"], case["n"])
    idmap = pp.paint_id_map(masks, spec.prompt_numbers, order=case["order"])
    assert hashlib.sha256(np.ascontiguousarray(idmap).tobytes()).hexdigest() == case["sha256"]
    boxes = pp.region_boxes(spec, masks, reference_quirks=True)
    assert {k: list(map(float, v)) for k, v in boxes.items()} == {k: v for k, v in case["bboxes"].items()}
    assert len(set(map(tuple, case["bboxes"].values()))) == 1          # the quirk itself: one box for all regions
    clean = pp.region_boxes(spec, masks)
    assert len(set(clean.values())) == case["n"]
    img = np.zeros((case["h"], case["w"], 3), np.uint8)
    req = pp.build_request(spec, img, masks, text_ids=list(range(100, 140)), max_num_tiles=4, dtype=torch.float32,
                           reference_quirks=True, prompt_order=case["order"])
    ids = req["input_ids"][0]
    starts = [int((ids == spec.crop_tokens_ids[k]).nonzero().min()) for k in case["order"]]
    assert starts == sorted(starts)                                     # crop runs listed in the reference's order
    assert {k: list(map(float, v)) for k, v in req["bboxes"][0].items()} == case["bboxes"]


# ---------------------------------------------------------------------------------------------------------------
# N1 device half, host tables: the int16 antialias weight tables the GPU resize passes consume reproduce torchvision's
# uint8 bicubic resize bit for bit when applied with plain integer arithmetic (horizontal pass first)
# ---------------------------------------------------------------------------------------------------------------
def _resize_with_tables(img_chw, oh, ow):
    from grasp_any_region_b200.preprocess_device import aa_tables

    def one_axis(a, out_size, axis):
        if a.shape[axis] == out_size:
            return a
        lo, size, w, prec = aa_tables(a.shape[axis], out_size)
        a = np.moveaxis(a, axis, -1).astype(np.int64)
        out = np.zeros(a.shape[:-1] + (out_size,), np.int64)
        for i in range(out_size):
            out[..., i] = (1 << (prec - 1)) + (a[..., lo[i]:lo[i] + size[i]] * w[i, :size[i]].astype(np.int64)).sum(-1)
        return np.moveaxis(np.clip(out >> prec, 0, 255).astype(np.uint8), -1, axis)

    return one_axis(one_axis(img_chw, ow, 2), oh, 1)


@pytest.mark.parametrize("h,w,oh,ow", [(100, 130, 56, 56), (64, 64, 112, 168), (300, 200, 448, 448), (1024, 1024, 448, 448),
                                       (200, 333, 224, 112), (77, 500, 448, 896), (448, 300, 448, 448), (61, 61, 61, 61)])
def test_aa_tables_reproduce_torchvision_uint8_bicubic(h, w, oh, ow):
    from torchvision.transforms import InterpolationMode
    from torchvision.transforms.v2 import functional as F

    g = torch.Generator().manual_seed(h * 31 + w)
    img = torch.randint(0, 256, (3, h, w), generator=g, dtype=torch.uint8)
    want = F.resize(img[None], (oh, ow), interpolation=InterpolationMode.BICUBIC)[0].numpy()
    assert np.array_equal(_resize_with_tables(img.numpy(), oh, ow), want)


def test_nearest_table_is_the_host_builders_expression():
    from grasp_any_region_b200.preprocess_device import nearest_table

    for n_in, n_out in ((1024, 448), (600, 1792), (333, 448), (448, 448), (5, 448)):
        a = torch.arange(n_in * 3).reshape(n_in, 3)
        assert torch.equal(pp._nearest(a, n_out, 3)[:, 0] // 3, nearest_table(n_in, n_out).long())
