"""Request preprocessing on the device: the uint8 image (3 MB for 1024^2) and the uint8 region-id map
(1 MB) cross PCIe; the bf16 `pixel_values` / `global_mask_values` tile tensors (41 MB per 17-tile request) are produced on
the GPU, bit-identical to the host builder (`preprocess.tile_image` / `tile_id_map`, which call the same torchvision
functional the reference's fast image processor calls, image_processing_perception_lm_fast.py:254-372).

What runs where:
  host   : the canvas choice (`preprocess.select_canvas`), the id-map painting and boxes (`preprocess.paint_id_map`,
           `region_boxes`), and the small per-axis tables -- ATen's int16 fixed-point antialias weights
           (`aa_tables`, a restatement of `_compute_index_ranges_int16_weights` in double precision) and torch's own
           NEAREST index expression (`nearest_table`);
  device : the two integer resize passes (horizontal first, uint8 intermediate), `_split` + rescale + normalise into
           tiles, the NEAREST id-map tiling (csrc/preprocess_u8.cu through the C ABI).
No CPU fallback: without the library or a CUDA tensor the calls raise.
"""
from __future__ import annotations

import ctypes as C
import math
from functools import lru_cache
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check
from .config import GARSpec
from .preprocess import paint_id_map, region_boxes, select_canvas


def aa_tables(in_size: int, out_size: int):
    """(idx_min int32[out], idx_size int32[out], weights int16[out, max_interp], precision) of ATen's uint8 antialiased
    bicubic resize along one axis (align_corners=False, scale = in / out):  cubic convolution with a = -0.5 over
    support 2 * max(scale, 1); weights normalised to sum 1, then scaled by 2^precision where precision is the largest
    value (< 22) that keeps the largest weight inside int16; round half away from zero."""
    scale = in_size / out_size
    support = 2.0 * scale if scale >= 1.0 else 2.0
    max_interp = int(math.ceil(support)) * 2 + 1
    invscale = 1.0 / scale if scale >= 1.0 else 1.0
    i = np.arange(out_size, dtype=np.float64)
    center = scale * (i + 0.5)
    lo = np.maximum((center - support + 0.5).astype(np.int64), 0)
    size = np.minimum((center + support + 0.5).astype(np.int64), in_size) - lo
    size = np.clip(size, 0, max_interp)
    j = np.arange(max_interp, dtype=np.float64)[None, :]
    x = np.abs((j + lo[:, None] - center[:, None] + 0.5) * invscale)
    a = -0.5
    w = np.where(x < 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0,
                 np.where(x < 2.0, ((a * x - 5.0 * a) * x + 8.0 * a) * x - 4.0 * a, 0.0))
    w = np.where(j < size[:, None], w, 0.0)
    # ATen accumulates the total left to right in double; np.cumsum does the same
    total = np.cumsum(w, axis=1)[:, -1]
    w = np.where(total[:, None] != 0.0, w / np.where(total == 0.0, 1.0, total)[:, None], w)
    w = np.where(j < size[:, None], w, 0.0)
    wt_max = float(w.max()) if w.size else 0.0
    precision = 0
    while precision < 22:
        if int(0.5 + wt_max * (1 << (precision + 1))) >= (1 << 15):
            break
        precision += 1
    scaled = w * float(1 << precision)
    wi = np.where(w < 0, (-0.5 + scaled).astype(np.int64), (0.5 + scaled).astype(np.int64)).astype(np.int16)
    return lo.astype(np.int32), size.astype(np.int32), wi, precision


def nearest_table(in_size: int, out_size: int) -> torch.Tensor:
    """torch's own expression in `preprocess._nearest` (fp32 arithmetic): floor((i + 0.5) * in / out), clamped"""
    return torch.clamp(((torch.arange(out_size) + 0.5) * in_size / out_size).long(), max=in_size - 1).to(torch.int32)


class DevicePreprocessor:
    """uint8 image (3, H, W) + uint8 id map (H, W), both CUDA tensors -> (pixel_values, global_mask_values, (ncw, nch))."""

    def __init__(self, spec: GARSpec, device="cuda:0"):
        if not torch.cuda.is_available():
            raise _lib.GarError("grasp_any_region_b200 needs a CUDA device: there is no CPU fallback")
        self.lib = _lib.load()
        self.spec = spec
        self.device = torch.device(device)
        self.tile = spec.vision.img_size
        self._aa: Dict[Tuple[int, int], tuple] = {}
        self._nn: Dict[Tuple[int, int], torch.Tensor] = {}

    def _aa_dev(self, in_size, out_size):
        key = (in_size, out_size)
        if key not in self._aa:
            lo, size, w, prec = aa_tables(in_size, out_size)
            self._aa[key] = (torch.from_numpy(lo).to(self.device), torch.from_numpy(size).to(self.device),
                             torch.from_numpy(w).to(self.device).contiguous(), int(w.shape[1]), prec)
        return self._aa[key]

    def _nn_dev(self, in_size, out_size):
        key = (in_size, out_size)
        if key not in self._nn:
            self._nn[key] = nearest_table(in_size, out_size).to(self.device)
        return self._nn[key]

    def _resize(self, img: torch.Tensor, oh: int, ow: int, s) -> torch.Tensor:
        """ATen order: horizontal pass, uint8 intermediate, vertical pass; a pass whose size does not change is skipped"""
        c, h, w = img.shape
        cur = img
        if ow != w:
            lo, size, wt, taps, prec = self._aa_dev(w, ow)
            out = torch.empty((c, h, ow), dtype=torch.uint8, device=self.device)
            check(self.lib.gar_resize_aa_u8_pass(C.c_void_p(cur.data_ptr()), c, h, w, 1, ow, C.c_void_p(lo.data_ptr()),
                                                 C.c_void_p(size.data_ptr()), C.c_void_p(wt.data_ptr()), taps, prec,
                                                 C.c_void_p(out.data_ptr()), s), "gar_resize_aa_u8_pass")
            cur, w = out, ow
        if oh != h:
            lo, size, wt, taps, prec = self._aa_dev(h, oh)
            out = torch.empty((c, oh, w), dtype=torch.uint8, device=self.device)
            check(self.lib.gar_resize_aa_u8_pass(C.c_void_p(cur.data_ptr()), c, h, w, 0, oh, C.c_void_p(lo.data_ptr()),
                                                 C.c_void_p(size.data_ptr()), C.c_void_p(wt.data_ptr()), taps, prec,
                                                 C.c_void_p(out.data_ptr()), s), "gar_resize_aa_u8_pass")
            cur = out
        return cur

    def __call__(self, image_chw_u8: torch.Tensor, idmap_hw_u8: Optional[torch.Tensor], max_num_tiles: int = 16,
                 stream: Optional[torch.cuda.Stream] = None):
        if not image_chw_u8.is_cuda or image_chw_u8.dtype != torch.uint8 or image_chw_u8.dim() != 3:
            raise _lib.GarError("DevicePreprocessor needs a CUDA uint8 (3, H, W) image (no CPU path)")
        image_chw_u8 = image_chw_u8.contiguous()
        t = self.tile
        _, h, w = image_chw_u8.shape
        ncw, nch = select_canvas(w, h, t, max_num_tiles)
        T = 1 + ncw * nch
        st = stream or torch.cuda.current_stream(self.device)
        s = C.c_void_p(st.cuda_stream)
        with torch.cuda.stream(st):
            px = torch.empty((T, 3, t, t), dtype=torch.bfloat16, device=self.device)
            thumb = self._resize(image_chw_u8, t, t, s)
            canvas = self._resize(image_chw_u8, nch * t, ncw * t, s)
            check(self.lib.gar_u8_to_tiles(C.c_void_p(thumb.data_ptr()), 1, 1, t, C.c_void_p(px.data_ptr()), s),
                  "gar_u8_to_tiles")
            check(self.lib.gar_u8_to_tiles(C.c_void_p(canvas.data_ptr()), ncw, nch, t, C.c_void_p(px[1:].data_ptr()), s),
                  "gar_u8_to_tiles")
            mv = None
            if idmap_hw_u8 is not None:
                if not idmap_hw_u8.is_cuda or idmap_hw_u8.dtype != torch.uint8 or tuple(idmap_hw_u8.shape) != (h, w):
                    raise _lib.GarError("id map must be a CUDA uint8 (H, W) tensor of the image's size")
                idmap_hw_u8 = idmap_hw_u8.contiguous()
                mv = torch.empty((T, 3, t, t), dtype=torch.bfloat16, device=self.device)
                tabs = [self._nn_dev(h, t), self._nn_dev(w, t), self._nn_dev(h, nch * t), self._nn_dev(w, ncw * t)]
                check(self.lib.gar_idmap_to_tiles(C.c_void_p(idmap_hw_u8.data_ptr()), h, w,
                                                  *[C.c_void_p(x.data_ptr()) for x in tabs], ncw, nch, t,
                                                  C.c_void_p(mv.data_ptr()), s), "gar_idmap_to_tiles")
            for x in (image_chw_u8, idmap_hw_u8, thumb, canvas):
                if x is not None:
                    x.record_stream(st)
        return px, mv, (ncw, nch)


def build_raw_request(spec: GARSpec, image_hwc_u8: np.ndarray, masks: Sequence[np.ndarray], text_ids: Sequence[int],
                      max_num_tiles: int = 16, prompt_ids: Optional[Sequence[int]] = None,
                      reference_quirks: bool = False, prompt_order: Optional[Sequence[int]] = None,
                      pin: bool = True) -> Dict:
    """The host half of a request for the device path: uint8 image (3, H, W) and uint8 id map (H, W) in (pinned) host
    memory, prompt ids, boxes -- everything `GARModel.generate_batch` needs to finish the request on the GPU.  Same
    prompt layout and options as `preprocess.build_request`."""
    from .synth import BOS, EH, EOT, SH

    tile = spec.vision.img_size
    prompt_ids = list(range(len(masks))) if prompt_ids is None else [int(k) for k in prompt_ids]
    order = list(range(len(masks))) if prompt_order is None else [int(i) for i in prompt_order]
    h, w = image_hwc_u8.shape[:2]
    ncw, nch = select_canvas(w, h, tile, max_num_tiles)
    T = 1 + ncw * nch
    img = torch.from_numpy(np.ascontiguousarray(image_hwc_u8)).permute(2, 0, 1).contiguous()
    idmap = torch.from_numpy(paint_id_map(masks, spec.prompt_numbers, prompt_ids, order if reference_quirks else None))
    text = [int(t) for t in text_ids]
    per = max(1, len(text) // (len(masks) + 1))
    seq = [BOS, SH, 882, EH, 271] + [spec.image_token_id] * (T * spec.tokens_per_tile)
    for n, i in enumerate(order if reference_quirks else range(len(masks))):
        seq += text[n * per:(n + 1) * per] + [spec.crop_tokens_ids[prompt_ids[i]]] * spec.tokens_per_tile
    seq += text[len(masks) * per:] + [EOT, SH, 78191, EH, 271]
    input_ids = torch.tensor(seq, dtype=torch.long)[None]
    if pin and torch.cuda.is_available():
        img, idmap, input_ids = img.pin_memory(), idmap.pin_memory(), input_ids.pin_memory()
    return dict(image_u8=img, idmap_u8=idmap, input_ids=input_ids, max_num_tiles=int(max_num_tiles),
                bboxes=[region_boxes(spec, masks, prompt_ids, reference_quirks)],
                aspect_ratios=torch.tensor([[ncw, nch]], dtype=torch.long))
