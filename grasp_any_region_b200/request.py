"""Request builders with a tokenizer: image + region mask(s) (+ question) -> the dict `GARModel.generate(**request)`
takes, without the reference's un-vendored hub processor.

Mirrors, text for text, what the reference's datasets feed the processor:
  * `SingleRegionCaptionDataset.__getitem__` (evaluation/eval_dataset.py:101-149): the caption question, one 256-long run
    of the region's crop token, the chat template with the generation prompt;
  * `MultiRegionDataset` (:192-313): the object list / per-object crop runs / question layout -- including, with
    reference_quirks=True, its stale-last-mask boxes and set-ordered regions (quirk 6);
  * the processor's media-token expansion (models/modeling/processing_perception_lm.py:200-220): every `<|image|>` becomes
    tiles * (tile // patch // pooling)^2 image tokens; then a plain tokenizer call.
Pixel / id-map tensors come from `preprocess` (host, the same torchvision resize the reference's image processor calls) or,
with raw=True, stay uint8 for the device path (`preprocess_device`).  `tokenizer` is any HF tokenizer of the checkpoint
(AutoTokenizer.from_pretrained(dir)); nothing here needs the network.
"""
from __future__ import annotations

import re
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from .config import GARSpec
from .preprocess import paint_id_map, region_boxes, select_canvas, tile_id_map, tile_image

IMAGE_TOKEN = "<|image|>"


def _crop_token(prompt_idx: int) -> str:
    return f"<|reserved_special_token_{prompt_idx + 2}|>"  # "actually starts from 2" (eval_dataset.py:107)


def _tokenize(spec: GARSpec, tokenizer, question: str, n_tiles: int) -> torch.Tensor:
    messages = [{"role": "user", "content": [{"type": "image"}, {"type": "text", "text": question}]}]
    raw_prompt = tokenizer.apply_chat_template(messages, add_generation_prompt=True, tokenize=False)
    image_token = getattr(tokenizer, "image_token", None) or IMAGE_TOKEN
    if raw_prompt.count(image_token) != 1:
        raise ValueError(f"the chat template must emit exactly one {image_token} for the image, got "
                         f"{raw_prompt.count(image_token)}")
    raw_prompt = raw_prompt.replace(image_token, image_token * (n_tiles * spec.tokens_per_tile))
    ids = tokenizer([raw_prompt])["input_ids"][0]
    return torch.tensor(ids, dtype=torch.long)[None]


def _finish(spec, tokenizer, image_hwc_u8, idmap, bboxes, question, max_num_tiles, dtype, device, raw) -> Dict:
    tile = spec.vision.img_size
    h, w = image_hwc_u8.shape[:2]
    ncw, nch = select_canvas(w, h, tile, max_num_tiles)
    input_ids = _tokenize(spec, tokenizer, question, 1 + ncw * nch)
    img = torch.from_numpy(np.ascontiguousarray(image_hwc_u8)).permute(2, 0, 1).contiguous()
    out = dict(input_ids=input_ids, bboxes=[bboxes], aspect_ratios=torch.tensor([[ncw, nch]], dtype=torch.long))
    if raw:
        out.update(image_u8=img, idmap_u8=torch.from_numpy(idmap), max_num_tiles=int(max_num_tiles))
        return out
    pixel_values, _ = tile_image(img, tile, max_num_tiles)
    out.update(pixel_values=pixel_values.to(dtype).to(device),
               global_mask_values=tile_id_map(torch.from_numpy(idmap), ncw, nch, tile).to(dtype).to(device),
               attention_mask=torch.ones_like(input_ids).to(dtype).to(device))  # bf16 ones, quirk 8
    out["input_ids"] = input_ids.to(device)
    out["aspect_ratios"] = out["aspect_ratios"].to(device)
    return out


def single_region_request(spec: GARSpec, tokenizer, image_hwc_u8: np.ndarray, mask_hw: np.ndarray,
                          prompt_token: str = "<Prompt0>", max_num_tiles: int = 36, dtype=torch.bfloat16,
                          device: str = "cpu", raw: bool = False) -> Dict:
    """`SingleRegionCaptionDataset(image, mask, processor, prompt_token)[0]` (demo/gar_with_mask.py:96-110)."""
    k = int(re.match(r"<Prompt(\d+)>", prompt_token).group(1))
    qs = (f"There are some objects I am curious about: {prompt_token};\n{prompt_token}: "
          f"{_crop_token(k) * spec.tokens_per_tile}Describe this masked region in detail.")
    idmap = paint_id_map([mask_hw], spec.prompt_numbers, prompt_ids=[k])
    bboxes = region_boxes(spec, [mask_hw], prompt_ids=[k])
    return _finish(spec, tokenizer, image_hwc_u8, idmap, bboxes, qs, max_num_tiles, dtype, device, raw)


def multi_region_request(spec: GARSpec, tokenizer, image_hwc_u8: np.ndarray, masks: Sequence[np.ndarray],
                         question_str: str, max_num_tiles: int = 36, dtype=torch.bfloat16, device: str = "cpu",
                         raw: bool = False, reference_quirks: bool = True,
                         prompt_order: Optional[Sequence[int]] = None) -> Dict:
    """`MultiRegionDataset(image, masks, question_str, processor)[0]` (demo/gar_relationship.py:100-125).  With
    reference_quirks=True (default: be the drop-in) every region's box is the LAST mask's box and the regions are visited
    in the order of a `set` of the question's <PromptK> strings, exactly as the reference does in this process
    (`prompt_order` overrides that order with one observed elsewhere); False gives each region its own box, in index order."""
    found = re.findall(r"<Prompt\d+>", question_str)
    matches = list(set(found)) if reference_quirks else sorted(set(found), key=lambda s: int(s[7:-1]))
    assert len(matches) == len(masks)
    if prompt_order is not None:
        matches = [f"<Prompt{int(i)}>" for i in prompt_order]
    order = [int(re.match(r"<Prompt(\d+)>", m).group(1)) for m in matches]
    objects_desc, sub_image_desc = "There are some objects I am curious about: ", ""
    for m, k in zip(matches, order):
        objects_desc += f"{m}; "
        sub_image_desc += f"{m}: {_crop_token(k) * spec.tokens_per_tile}\n"
    prompt = objects_desc + "\n" + sub_image_desc + "\n" + question_str
    idmap = paint_id_map(masks, spec.prompt_numbers, order=order if reference_quirks else None)
    bboxes = region_boxes(spec, masks, reference_quirks=reference_quirks)
    return _finish(spec, tokenizer, image_hwc_u8, idmap, bboxes, prompt, max_num_tiles, dtype, device, raw)
