"""GAR-Bench driver as a regression suite: replays the reference's benchmark loop
(evaluation/GAR-Bench/inference.py:114-203) on the B200 `GARModel` and compares every answer, string for string, with the
`model_outputs/*.json` caches the reference ships (README.md:79,102,131) -- the only externally pinned outputs of the path.

Needs what the reference needs (a GAR checkpoint directory with its tokenizer, the benchmark images); nothing here needs the
network or the packages this image lacks: COCO RLE masks are decoded by `decode_rle` (pycocotools is not installed), requests
are built by `request.multi_region_request` with reference_quirks=True (the caches were produced THROUGH the reference's
stale-box builder, so reproducing them needs the quirks).

    python -m grasp_any_region_b200.gar_bench --model_name_or_path <dir> --image_folder evaluation/GAR-Bench/annotations \\
        --cache evaluation/GAR-Bench/model_outputs/gar_1b_vqa.json --mode vqa
"""
from __future__ import annotations

import argparse
import json
import os
from typing import Callable, Dict, List, Sequence

import numpy as np


def decode_rle(rle: Dict) -> np.ndarray:
    """COCO run-length mask -> (h, w) uint8.  `counts` is either a list of run lengths or pycocotools' compressed string
    (6 bits per character from chr(48), bit 5 = continuation, sign-extended, delta-coded against the run two back);
    runs alternate 0 / 1 starting with 0 over the column-major (Fortran-order) pixels."""
    h, w = int(rle["size"][0]), int(rle["size"][1])
    counts = rle["counts"]
    if isinstance(counts, (bytes, str)):
        s = counts.decode("ascii") if isinstance(counts, bytes) else counts
        runs: List[int] = []
        p = 0
        while p < len(s):
            x, k, more = 0, 0, True
            while more:
                c = ord(s[p]) - 48
                x |= (c & 0x1F) << (5 * k)
                more = bool(c & 0x20)
                p += 1
                k += 1
                if not more and (c & 0x10):
                    x |= -1 << (5 * k)
            if len(runs) > 2:
                x += runs[-2]
            runs.append(x)
    else:
        runs = [int(c) for c in counts]
    if sum(runs) != h * w or any(r < 0 for r in runs):
        raise ValueError(f"corrupt RLE: runs sum to {sum(runs)}, mask has {h * w} pixels")
    flat = np.zeros(h * w, dtype=np.uint8)
    pos, val = 0, 0
    for r in runs:
        if val:
            flat[pos:pos + r] = 1
        pos += r
        val ^= 1
    return flat.reshape((w, h)).T.copy()


def question_for(item: Dict, mode: str) -> str:
    """evaluation/GAR-Bench/inference.py:121-131"""
    if mode == "vqa":
        q = f"Question: {item['question']}\nOptions:"
        for op in item["choices"]:
            q += f"\n{op}"
        return q + "\nAnswer with the correct option's letter directly."
    if mode == "simple":
        return item["question"]
    if mode == "detailed":
        return "Describe <Prompt0> in detail, including the relationship with <Prompt1>."
    raise NotImplementedError(mode)


def clean_output(text: str) -> str:
    """:170-175 -- decoded with special tokens kept, stripped, the end-of-turn marker removed"""
    text = text.strip()
    if text.endswith("<|eot_id|>"):
        text = text.replace("<|eot_id|>", "")
    return text


def replay(items: Sequence[Dict], mode: str, answer_fn: Callable[[Dict, str, List[np.ndarray]], str]) -> Dict:
    """Runs `answer_fn(item, question, masks)` over cached benchmark items and compares with the cached `model_output`.
    Returns per-item records and the summary the reference prints (exact-match accuracy per type for vqa, :185-203)."""
    records = []
    for item in items:
        masks = [decode_rle(r) for r in item["mask_rles"]]
        out = clean_output(answer_fn(item, question_for(item, mode), masks))
        records.append({"image": item.get("image"), "type": item.get("type"), "answer": item.get("answer"),
                        "cached": item.get("model_output"), "ours": out, "same": out == item.get("model_output")})
    summary = {"items": len(records), "identical_to_cache": sum(r["same"] for r in records)}
    if mode == "vqa":
        acc = lambda rs, key: sum(str(r[key]).lower() == str(r["answer"]).lower() for r in rs)  # noqa: E731
        summary["accuracy_ours"] = acc(records, "ours") / max(1, len(records))
        summary["accuracy_cached"] = acc(records, "cached") / max(1, len(records))
        summary["per_type"] = {t: {"ours": acc([r for r in records if r["type"] == t], "ours"),
                                   "cached": acc([r for r in records if r["type"] == t], "cached"),
                                   "total": sum(r["type"] == t for r in records)}
                               for t in sorted({r["type"] for r in records})}
    return {"summary": summary, "records": records}


def main():  # pragma: no cover  (needs a checkpoint, its tokenizer and the benchmark images)
    import torch
    from PIL import Image
    from transformers import AutoTokenizer

    from . import GARModel
    from .request import multi_region_request

    ap = argparse.ArgumentParser()
    ap.add_argument("--model_name_or_path", required=True)
    ap.add_argument("--image_folder", required=True)
    ap.add_argument("--cache", required=True, help="one of the reference's evaluation/GAR-Bench/model_outputs/*.json")
    ap.add_argument("--mode", required=True, choices=["vqa", "simple", "detailed"])
    ap.add_argument("--max_num_tiles", type=int, default=36)
    ap.add_argument("--limit", type=int, default=0)
    args = ap.parse_args()
    model = GARModel.from_pretrained(args.model_name_or_path, torch_dtype=torch.bfloat16, device_map="cuda:0").eval()
    tok = AutoTokenizer.from_pretrained(args.model_name_or_path)
    model.configure_engine(max_new_tokens=1024, max_batch=1)

    def answer(item, question, masks):
        img = np.array(Image.open(os.path.join(args.image_folder, item["image"])).convert("RGB"))
        req = multi_region_request(model.spec, tok, img, [m.astype(bool) for m in masks], question,
                                   max_num_tiles=args.max_num_tiles, raw=True, reference_quirks=True)
        ids = model.generate_batch([req], max_new_tokens=1024, eos_token_id=tok.eos_token_id)[0]
        return tok.decode(ids, skip_special_tokens=False)

    items = json.load(open(args.cache))
    res = replay(items[: args.limit] if args.limit else items, args.mode, answer)
    print(json.dumps(res["summary"], indent=1))
    for r in res["records"]:
        if not r["same"]:
            print("DIFF", r["image"], "| cached:", r["cached"][:80], "| ours:", r["ours"][:80])


if __name__ == "__main__":  # pragma: no cover
    main()
