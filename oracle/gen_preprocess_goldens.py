"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/preprocess_goldens.json from the reference's own tiler.

Runs in the build container only (needs /root/reference).  The reference's image processor module does not import
under transformers 5.5 (`DefaultFastImageProcessorKwargs` is gone), so its pure-Python canvas functions and `_split`
are lifted *by source text* out of
  projects/grasp_any_region/models/modeling/image_processing_perception_lm_fast.py:95-252
into a stub class and executed unmodified; `resize` is driven exactly as `_preprocess` (:296-372) drives it
(thumbnail with max_num_tiles=1, then the tiled canvas, `_split`, thumbnail first, rescale 1/255, normalise 0.5/0.5).
"""
import ast
import hashlib
import json
import math
import os
import sys
import textwrap
from functools import reduce

import numpy as np
import torch
from torchvision.transforms.v2 import functional as F
from torchvision.transforms import InterpolationMode

REF = "/root/reference/projects/grasp_any_region/models/modeling/image_processing_perception_lm_fast.py"
WANT = ["_factors", "_find_supported_aspect_ratios", "_get_image_height_width", "_fit_image_to_canvas",
        "_find_closest_aspect_ratio", "_split"]


def lift():
    src = open(REF).read()
    tree = ast.parse(src)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "PerceptionLMImageProcessorFast"][0]
    parts = []
    for n in cls.body:
        if isinstance(n, ast.FunctionDef) and n.name in WANT:
            seg = ast.get_source_segment(src, n)
            if n.name == "_factors":
                seg = "@staticmethod\n" + seg
            parts.append(textwrap.indent(seg, "    "))
    code = "class Ref:\n    def __init__(self, max_num_tiles):\n        self.max_num_tiles = max_num_tiles\n\n" + "\n\n".join(parts)
    ns = {"math": math, "reduce": reduce, "torch": torch}
    exec(code, ns)
    return ns["Ref"]


def ref_canvas(ref, w, h, tile):
    """image_processing_perception_lm_fast.py:254-277 (`resize`), the canvas decision only"""
    ar = ref._fit_image_to_canvas(img_width=w, img_height=h, tile_size=tile)
    if ar is None:
        ar = ref._find_closest_aspect_ratio(img_width=w, img_height=h, tile_size=tile)
    return ar


def ref_pixels(ref, img_chw_u8, tile):
    thumb = F.resize(img_chw_u8[None], (tile, tile), interpolation=InterpolationMode.BICUBIC)
    nw, nh = ref_canvas(ref, img_chw_u8.shape[2], img_chw_u8.shape[1], tile)
    big = F.resize(img_chw_u8[None], (nh * tile, nw * tile), interpolation=InterpolationMode.BICUBIC)
    tiles = ref._split(big, nw, nh)
    stacked = torch.cat([thumb.unsqueeze(1), tiles], dim=1)[0]
    x = stacked.float() * (1.0 / 255.0)
    return (x - 0.5) / 0.5, (nw, nh)


def digest(t):
    return hashlib.sha256(np.ascontiguousarray(t.numpy()).tobytes()).hexdigest()


def main():
    Ref = lift()
    out = {"canvas": [], "pixels": []}
    rng = np.random.default_rng(0)
    sizes = [(1024, 1024), (448, 448), (449, 448), (100, 2000), (2000, 100), (1920, 1080), (1080, 1920), (640, 480),
             (3000, 2000), (5000, 300), (333, 777), (896, 896), (897, 896), (1344, 448), (300, 300), (1, 1), (7000, 7000)]
    sizes += [(int(a), int(b)) for a, b in rng.integers(16, 4000, size=(60, 2))]
    for mt in (1, 4, 8, 16, 36):
        ref = Ref(mt)
        for (w, h) in sizes:
            ar = ref_canvas(ref, w, h, 448) if mt > 1 else (1, 1)
            out["canvas"].append({"w": w, "h": h, "max_num_tiles": mt, "ncw": ar[0], "nch": ar[1]})
    for (w, h, mt, tile) in ((300, 500, 8, 56), (500, 300, 8, 56), (200, 200, 16, 28), (90, 700, 16, 28)):
        g = torch.Generator().manual_seed(w * 7 + h)
        img = torch.randint(0, 256, (3, h, w), generator=g, dtype=torch.uint8)
        px, (nw, nh) = ref_pixels(Ref(mt), img, tile)
        out["pixels"].append({"w": w, "h": h, "max_num_tiles": mt, "tile": tile, "seed": w * 7 + h, "ncw": nw, "nch": nh,
                              "shape": list(px.shape), "sha256": digest(px), "sum": float(px.double().sum()),
                              "probe": [float(px[t, c, y, x]) for (t, c, y, x) in ((0, 0, 0, 0), (1, 1, 3, 5), (px.shape[0] - 1, 2, tile - 1, tile - 1))]})
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "preprocess_goldens.json")
    json.dump(out, open(dst, "w"))
    print("wrote", dst, len(out["canvas"]), "canvas cases,", len(out["pixels"]), "pixel cases")


if __name__ == "__main__":
    main()
