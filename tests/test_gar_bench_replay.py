"""SURVEY §8f N3 scaffolding on CPU: the GAR-Bench replay driver (grasp_any_region_b200/gar_bench.py) decodes the
benchmark's COCO RLE masks without pycocotools, rebuilds the reference driver's questions
(evaluation/GAR-Bench/inference.py:121-131), post-processes answers like it (:170-175) and scores them against the cached
`model_outputs/*.json` the reference ships.  The fixture holds five real items of those caches."""
import json
import os

import numpy as np

from grasp_any_region_b200.gar_bench import clean_output, decode_rle, question_for, replay

ITEMS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "gar_bench_items.json")))


def _encode_rle(mask):
    """pycocotools' compressed RLE, restated for the round-trip check (column-major runs, 5-bit groups, delta vs run-2)"""
    h, w = mask.shape
    flat = mask.T.reshape(-1)
    runs, cur, n = [], 0, 0
    for v in flat:
        if v != cur:
            runs.append(n)
            cur, n = v, 0
        n += 1
    runs.append(n)
    out = []
    for i, r in enumerate(runs):
        x = r - runs[i - 2] if i > 2 else r
        more = True
        while more:
            c = x & 0x1F
            x >>= 5
            more = not (x == 0 and not (c & 0x10)) and not (x == -1 and (c & 0x10))
            if more:
                c |= 0x20
            out.append(chr(c + 48))
    return {"size": [h, w], "counts": "".join(out)}


def test_rle_round_trip_and_real_vectors():
    rng = np.random.default_rng(0)
    for h, w in ((7, 5), (64, 48), (300, 431)):
        m = np.zeros((h, w), np.uint8)
        for _ in range(4):
            y, x = int(rng.integers(0, h - 2)), int(rng.integers(0, w - 2))
            m[y: y + int(rng.integers(1, h)), x: x + int(rng.integers(1, w))] = 1
        assert np.array_equal(decode_rle(_encode_rle(m)), m)
        assert np.array_equal(decode_rle({"size": [h, w], "counts": [h * w]}), np.zeros((h, w), np.uint8))
    for item in ITEMS["vqa"] + ITEMS["simple"]:
        for rle in item["mask_rles"]:
            m = decode_rle(rle)                         # raises unless the runs cover the mask exactly
            assert m.shape == tuple(rle["size"]) and 0 < m.sum() < m.size
            assert np.array_equal(decode_rle(_encode_rle(m)), m) and _encode_rle(m)["counts"] == rle["counts"]


def test_questions_and_scoring_follow_the_reference_driver():
    it = ITEMS["vqa"][0]
    q = question_for(it, "vqa")
    assert q.startswith("Question: " + it["question"] + "\nOptions:\n" + it["choices"][0])
    assert q.endswith("\nAnswer with the correct option's letter directly.")
    assert question_for(ITEMS["simple"][0], "simple") == ITEMS["simple"][0]["question"]
    assert clean_output("  A<|eot_id|> ") == "A" and clean_output("B") == "B"
    seen = []

    def echo(item, question, masks):
        seen.append((question, [m.shape for m in masks]))
        return item["model_output"] + "<|eot_id|>"

    res = replay(ITEMS["vqa"], "vqa", echo)
    assert res["summary"]["identical_to_cache"] == len(ITEMS["vqa"]) == res["summary"]["items"]
    assert res["summary"]["accuracy_ours"] == res["summary"]["accuracy_cached"]
    assert all(len(shapes) == len(item["mask_rles"]) for (_, shapes), item in zip(seen, ITEMS["vqa"]))
    wrong = replay(ITEMS["vqa"], "vqa", lambda item, q, m: "Z")
    assert wrong["summary"]["identical_to_cache"] == 0 and wrong["summary"]["accuracy_ours"] == 0.0
