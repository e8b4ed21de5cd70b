"""Seeded synthetic (image, mask-set, prompt) requests with the exact dict layout the
reference's request builder emits (reference: evaluation/eval_dataset.py:101-149 for one
region, :152-313 for several; mask-image convention GraspAnyRegion_Dataset.py:123-128).

No tokenizer / hub processor exists offline, so text ids are uniform random; the structure
(image-token run of T*256, one 256-long crop-token run per region, bbox = inclusive max
pixel / size, id-map tiled NEAREST and normalised to 2*id/255-1) follows the reference.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch

from .config import GARSpec

BOS, SH, EH, EOT = 128000, 128006, 128007, 128009


def tile_grid_for(image_hw: Tuple[int, int], max_num_tiles: int, tile: int = 448) -> Tuple[int, int]:
    """Square images only (the synthetic benchmark): largest n x n grid with n*n <= max tiles
    whose canvas does not exceed the up-scaled image (reference canvas search
    image_processing_perception_lm_fast.py:171-225 gives 4x4 for 1024^2 @16, 6x6 @36, 2x2 @8)."""
    h, w = image_hw
    assert h == w, "synthetic requests use square images"
    n = int(np.floor(np.sqrt(max_num_tiles)))
    return n, n  # (tiles_w, tiles_h)


def _nearest_resize(a: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    h, w = a.shape
    ys = np.minimum(((np.arange(out_h) + 0.5) * h / out_h).astype(np.int64), h - 1)
    xs = np.minimum(((np.arange(out_w) + 0.5) * w / out_w).astype(np.int64), w - 1)
    return a[ys][:, xs]


def make_request(spec: GARSpec, seed: int = 0, n_masks: int = 1, image_size: int = 1024,
                 max_num_tiles: int = 16, text_tokens: int = 64,
                 dtype: torch.dtype = torch.bfloat16, device: str = "cpu") -> Dict:
    g = np.random.default_rng(seed)
    tg = torch.Generator().manual_seed(seed)
    tile = spec.vision.img_size
    tw, th = tile_grid_for((image_size, image_size), max_num_tiles, tile)
    T = tw * th + 1
    P = spec.prompt_numbers
    assert 1 <= n_masks <= P

    # ---- id map: P everywhere, region k -> k, first writer wins (eval_dataset.py:238-239)
    idmap = np.full((image_size, image_size), P, dtype=np.uint8)
    bboxes: Dict[str, Tuple[float, float, float, float]] = {}
    for k in range(n_masks):
        for _ in range(64):
            side_w = int(g.integers(image_size // 16, 3 * image_size // 4))
            side_h = int(g.integers(image_size // 16, 3 * image_size // 4))
            x0 = int(g.integers(0, image_size - side_w))
            y0 = int(g.integers(0, image_size - side_h))
            m = np.zeros_like(idmap, dtype=bool)
            m[y0:y0 + side_h, x0:x0 + side_w] = True
            fill = m & (idmap == P)
            if fill.any():
                break
        idmap[fill] = k
        ys, xs = np.nonzero(m)
        bboxes[str(spec.crop_tokens_ids[k])] = (
            xs.min() / image_size, ys.min() / image_size, xs.max() / image_size, ys.max() / image_size)

    # ---- tiles: thumbnail first, then the grid row-major (tile 0 = thumbnail, quirk 3)
    canvas = _nearest_resize(idmap, th * tile, tw * tile)
    tiles = [_nearest_resize(idmap, tile, tile)]
    for r in range(th):
        for c in range(tw):
            tiles.append(canvas[r * tile:(r + 1) * tile, c * tile:(c + 1) * tile])
    ids = torch.from_numpy(np.stack(tiles)).float()  # (T,448,448)
    mask_values = (2.0 * ids / 255.0 - 1.0)[:, None].expand(T, 3, tile, tile).contiguous()

    pixel_values = torch.rand((T, 3, tile, tile), generator=tg) * 2.0 - 1.0

    # ---- prompt ids
    def rnd(n):
        return [int(v) for v in g.integers(0, 128000, size=n)]

    n_img = T * spec.tokens_per_tile
    per = max(text_tokens - 10, n_masks * 4) // n_masks
    seq: List[int] = [BOS, SH, 882, EH, 271] + [spec.image_token_id] * n_img
    for k in range(n_masks):
        seq += rnd(4) + [spec.crop_tokens_ids[k]] * spec.tokens_per_tile + rnd(per - 4)
    seq += [EOT, SH, 78191, EH, 271]
    input_ids = torch.tensor(seq, dtype=torch.long)[None]

    return dict(
        input_ids=input_ids.to(device),
        attention_mask=torch.ones_like(input_ids).to(dtype).to(device),
        pixel_values=pixel_values.to(dtype).to(device),
        global_mask_values=mask_values.to(dtype).to(device),
        bboxes=[bboxes],
        aspect_ratios=torch.tensor([[tw, th]], dtype=torch.long, device=device),
    )


def make_raw_request(spec: GARSpec, seed: int = 0, n_masks: int = 1, image_size: int = 1024, max_num_tiles: int = 16,
                     text_tokens: int = 64, pin: bool = True) -> Dict:
    """The same kind of request before preprocessing: a seeded uint8 image, rectangle masks and random text ids, as the
    host half of the device-preprocessing path (preprocess_device.build_raw_request).  Same prompt length as
    make_request for equal arguments."""
    from .preprocess_device import build_raw_request

    g = np.random.default_rng(seed)
    image = g.integers(0, 256, size=(image_size, image_size, 3), dtype=np.uint8)
    masks = []
    for _ in range(n_masks):
        side_w = int(g.integers(image_size // 16, 3 * image_size // 4))
        side_h = int(g.integers(image_size // 16, 3 * image_size // 4))
        x0 = int(g.integers(0, image_size - side_w))
        y0 = int(g.integers(0, image_size - side_h))
        m = np.zeros((image_size, image_size), dtype=bool)
        m[y0:y0 + side_h, x0:x0 + side_w] = True
        masks.append(m)
    n_text = n_masks * (max(text_tokens - 10, n_masks * 4) // n_masks)
    text = [int(v) for v in g.integers(0, 128000, size=n_text)]
    return build_raw_request(spec, image, masks, text, max_num_tiles=max_num_tiles, pin=pin)
