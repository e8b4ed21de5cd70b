"""Host-side request builder: image + region masks -> the tensors `GARModel.generate()` consumes (SURVEY §8 row R1,
next-row N1).  Restates, without the un-vendored hub processor:

  * the tile-canvas choice and tiling of the reference's image processor
    (models/modeling/image_processing_perception_lm_fast.py:95-277, 296-340): every grid (n_w, n_h) with
    n_w * n_h <= max_num_tiles is a candidate canvas; the image goes onto the canvas it fits without down-sampling
    that keeps its shorter edge largest (later candidates win ties), else onto the closest aspect ratio; it is
    resized to the WHOLE canvas (aspect ratio not preserved), split row-major into tiles, and a one-tile thumbnail
    goes in front; pixels are rescaled by 1/255 and normalised with mean = std = 0.5;
  * the region id map of the reference's request builders (evaluation/eval_dataset.py:56-99, 192-271): background =
    prompt_numbers, region k = k, first writer wins; boxes are (min, inclusive max) / size; the id map is tiled like
    the image but with NEAREST resampling, replicated to 3 channels and encoded as 2 v / 255 - 1 (the decode at
    modeling_gar.py:315-325 inverts it).

Resizing uses the same torchvision v2 functional the reference's fast processor calls, so pixel values are identical
by construction; `tests/test_preprocess.py` pins canvas choice, tile order and pixels to vectors generated from the
reference's own functions (`oracle/gen_preprocess_goldens.py`).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .config import GARSpec


def _factors(n: int) -> List[int]:
    out = set()
    for i in range(1, int(n ** 0.5) + 1):
        if n % i == 0:
            out.add(i)
            out.add(n // i)
    return sorted(out)


def supported_aspect_ratios(max_num_tiles: int) -> Dict[float, List[Tuple[int, int]]]:
    """aspect ratio -> grids (n_w, n_h), in the reference's insertion order (it decides ties)."""
    table: Dict[float, List[Tuple[int, int]]] = {}
    for chunks in range(max_num_tiles, 0, -1):
        for x in _factors(chunks):
            table.setdefault(x / (chunks // x), []).append((x, chunks // x))
    return table


def _resized_wh(img_w: int, img_h: int, target_w: int, target_h: int) -> Tuple[float, float]:
    """size the image would take on the canvas with its aspect ratio preserved (used only to rank canvases)"""
    scale = img_w / img_h
    factor = min(target_w / img_w, target_h / img_h)
    if scale > 1.0:
        new_w = factor * img_w
        return new_w, math.floor(new_w / scale)
    new_h = factor * img_h
    return math.floor(new_h * scale), new_h


def fit_image_to_canvas(img_w: int, img_h: int, tile: int, max_num_tiles: int) -> Optional[Tuple[int, int]]:
    best, best_wh = None, None
    scale = img_w / img_h
    for ratios in supported_aspect_ratios(max_num_tiles).values():
        for (n_w, n_h) in ratios:
            if n_w * tile < img_w or n_h * tile < img_h:
                continue
            wh = _resized_wh(img_w, img_h, n_w * tile, n_h * tile)
            if best is None or (scale < 1.0 and wh[0] >= best_wh[0]) or (scale >= 1.0 and wh[1] >= best_wh[1]):
                best, best_wh = (n_w, n_h), wh
    return best


def find_closest_aspect_ratio(img_w: int, img_h: int, max_num_tiles: int) -> Tuple[int, int]:
    target = img_w / img_h
    table = supported_aspect_ratios(max_num_tiles)
    if target >= 1:
        k = min((r for r in table if r <= target), key=lambda r: abs(r - target))
        return max(table[k], key=lambda g: g[0])
    k = min((r for r in table if r > target), key=lambda r: abs(1 / r - 1 / target))
    return max(table[k], key=lambda g: g[1])


def select_canvas(img_w: int, img_h: int, tile: int = 448, max_num_tiles: int = 16) -> Tuple[int, int]:
    """(tiles_w, tiles_h) = the `aspect_ratio` the reference's processor returns."""
    if max_num_tiles <= 1:
        return (1, 1)
    grid = fit_image_to_canvas(img_w, img_h, tile, max_num_tiles)
    return grid if grid is not None else find_closest_aspect_ratio(img_w, img_h, max_num_tiles)


def _split(canvas: torch.Tensor, ncw: int, nch: int) -> torch.Tensor:
    """(C, nch*t, ncw*t) -> (nch*ncw, C, t, t), tiles row-major over (row, column)"""
    c, h, w = canvas.shape
    return canvas.view(c, nch, h // nch, ncw, w // ncw).permute(1, 3, 0, 2, 4).reshape(nch * ncw, c, h // nch, w // ncw)


def tile_image(image_chw_u8: torch.Tensor, tile: int = 448, max_num_tiles: int = 16):
    """uint8 (3, H, W) -> pixel_values (1 + ncw*nch, 3, tile, tile) float32 in [-1, 1] (thumbnail first), (ncw, nch)"""
    from torchvision.transforms import InterpolationMode
    from torchvision.transforms.v2 import functional as F

    assert image_chw_u8.dtype == torch.uint8 and image_chw_u8.ndim == 3
    h, w = image_chw_u8.shape[1:]
    ncw, nch = select_canvas(w, h, tile, max_num_tiles)
    thumb = F.resize(image_chw_u8[None], (tile, tile), interpolation=InterpolationMode.BICUBIC)[0]
    canvas = F.resize(image_chw_u8[None], (nch * tile, ncw * tile), interpolation=InterpolationMode.BICUBIC)[0]
    tiles = torch.cat([thumb[None], _split(canvas, ncw, nch)], 0)
    return (tiles.float() * (1.0 / 255.0) - 0.5) / 0.5, (ncw, nch)


def _nearest(a: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """PIL / torch 'nearest' index rule: source index = floor((i + 0.5) * in / out)"""
    h, w = a.shape
    ys = torch.clamp(((torch.arange(out_h) + 0.5) * h / out_h).long(), max=h - 1)
    xs = torch.clamp(((torch.arange(out_w) + 0.5) * w / out_w).long(), max=w - 1)
    return a[ys][:, xs]


def tile_id_map(idmap_hw_u8: torch.Tensor, ncw: int, nch: int, tile: int = 448) -> torch.Tensor:
    """id map (H, W) uint8 -> global_mask_values (1 + ncw*nch, 3, tile, tile) float32 = 2 id / 255 - 1, NEAREST"""
    thumb = _nearest(idmap_hw_u8, tile, tile)
    canvas = _nearest(idmap_hw_u8, nch * tile, ncw * tile)
    tiles = torch.cat([thumb[None, None], _split(canvas[None], ncw, nch)], 0).float()
    return (2.0 * tiles / 255.0 - 1.0).expand(-1, 3, -1, -1).contiguous()


def bbox_of(mask_hw: np.ndarray) -> Tuple[float, float, float, float]:
    """(x_min / W, y_min / H, x_max / W, y_max / H) with the INCLUSIVE max pixel (eval_dataset.py:77-85)"""
    ys, xs = np.nonzero(mask_hw)
    h, w = mask_hw.shape
    return (xs.min() / w, ys.min() / h, xs.max() / w, ys.max() / h)


def paint_id_map(masks: Sequence[np.ndarray], prompt_numbers: int, prompt_ids: Optional[Sequence[int]] = None,
                 order: Optional[Sequence[int]] = None) -> np.ndarray:
    """background = prompt_numbers; mask i paints id prompt_ids[i] (default i) where still background -- first writer
    wins (eval_dataset.py:62-69, 238-239).  `order`: the sequence in which the masks are visited (default 0..n-1; the
    reference's multi-region builder visits a `set` of prompt strings, so its order varies with PYTHONHASHSEED)."""
    assert 1 <= len(masks) <= prompt_numbers
    prompt_ids = list(range(len(masks))) if prompt_ids is None else [int(k) for k in prompt_ids]
    assert all(0 <= k < prompt_numbers for k in prompt_ids) and len(prompt_ids) == len(masks)
    idmap = np.full(np.asarray(masks[0]).shape, prompt_numbers, dtype=np.uint8)
    for i in (range(len(masks)) if order is None else order):
        idmap[(idmap == prompt_numbers) & np.asarray(masks[i]).astype(bool)] = prompt_ids[i]
    return idmap


def region_boxes(spec: GARSpec, masks: Sequence[np.ndarray], prompt_ids: Optional[Sequence[int]] = None,
                 reference_quirks: bool = False) -> Dict[str, Tuple[float, float, float, float]]:
    """bboxes dict keyed by str(crop token id) (modeling_gar.py:367).  reference_quirks=True reproduces what
    `MultiRegionDataset._parse_annotations` really emits (eval_dataset.py:201-205, 241): the box of
    EVERY region is computed from the last mask, because the loop reuses a stale `mask_id`."""
    prompt_ids = list(range(len(masks))) if prompt_ids is None else [int(k) for k in prompt_ids]
    out = {}
    for i, k in enumerate(prompt_ids):
        src = masks[-1] if reference_quirks else masks[i]
        out[str(spec.crop_tokens_ids[k])] = bbox_of(np.asarray(src))
    return out


def build_request(spec: GARSpec, image_hwc_u8: np.ndarray, masks: Sequence[np.ndarray], text_ids: Sequence[int],
                  max_num_tiles: int = 16, dtype: torch.dtype = torch.bfloat16, device: str = "cpu",
                  prompt_ids: Optional[Sequence[int]] = None, reference_quirks: bool = False,
                  prompt_order: Optional[Sequence[int]] = None) -> Dict:
    """The dict `GARModel.generate(**request)` takes, for one image and its region masks.  `text_ids` stands in for the
    tokenised question (no tokenizer ships with the repo): the prompt is
    [BOS, header] + <|image|> * (T * 256) + per region k: text + crop_token_k * 256 + text + [eot, assistant header].

    prompt_ids[i]: which <PromptK> mask i is (default i).  reference_quirks=True emits exactly what the reference's
    `gar_relationship.py` request builder would (stale last-mask boxes; regions painted and listed in `prompt_order`,
    the visiting order of its `set` of prompt strings -- pass the order observed there; default 0..n-1)."""
    from .synth import BOS, EH, EOT, SH

    tile = spec.vision.img_size
    prompt_ids = list(range(len(masks))) if prompt_ids is None else [int(k) for k in prompt_ids]
    order = list(range(len(masks))) if prompt_order is None else [int(i) for i in prompt_order]
    assert sorted(order) == list(range(len(masks)))
    img = torch.from_numpy(np.ascontiguousarray(image_hwc_u8)).permute(2, 0, 1).contiguous()
    pixel_values, (ncw, nch) = tile_image(img, tile, max_num_tiles)
    idmap = paint_id_map(masks, spec.prompt_numbers, prompt_ids, order if reference_quirks else None)
    mask_values = tile_id_map(torch.from_numpy(idmap), ncw, nch, tile)
    bboxes = region_boxes(spec, masks, prompt_ids, reference_quirks)
    T = pixel_values.shape[0]
    text = [int(t) for t in text_ids]
    per = max(1, len(text) // (len(masks) + 1))
    seq: List[int] = [BOS, SH, 882, EH, 271] + [spec.image_token_id] * (T * spec.tokens_per_tile)
    for n, i in enumerate(order if reference_quirks else range(len(masks))):
        seq += text[n * per:(n + 1) * per] + [spec.crop_tokens_ids[prompt_ids[i]]] * spec.tokens_per_tile
    seq += text[len(masks) * per:] + [EOT, SH, 78191, EH, 271]
    input_ids = torch.tensor(seq, dtype=torch.long)[None]
    return dict(
        input_ids=input_ids.to(device),
        attention_mask=torch.ones_like(input_ids).to(dtype).to(device),
        pixel_values=pixel_values.to(dtype).to(device),
        global_mask_values=mask_values.to(dtype).to(device),
        bboxes=[bboxes],
        aspect_ratios=torch.tensor([[ncw, nch]], dtype=torch.long, device=device),
    )
