"""Tokenizer-side request builders (grasp_any_region_b200/request.py) against the reference's datasets: the prompt text is
checked against the strings `evaluation/eval_dataset.py` builds (lifted and executed by oracle/gen_idmap_goldens.py for
the multi-region case: tests/golden/idmap_goldens.json carries the visiting order), with a stub tokenizer that records
what it is asked to tokenize."""
import json
import os
import re

import numpy as np
import torch

from grasp_any_region_b200.config import gar_tiny
from grasp_any_region_b200.request import multi_region_request, single_region_request

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "idmap_goldens.json")))


class StubTokenizer:
    """whitespace-free toy: special tokens -> their ids, every other character -> 1000 + ord"""
    image_token = "<|image|>"
    SPECIAL = {"<|image|>": 128002, "<|begin_of_text|>": 128000, "<|start_header_id|>": 128006, "<|end_header_id|>": 128007,
               "<|eot_id|>": 128009, **{f"<|reserved_special_token_{i}|>": v for i, v in
                                        {2: 128004, 3: 128005, 4: 128008, 5: 128010, 6: 128011}.items()},
               **{f"<Prompt{i}>": 128256 + i for i in range(5)}}

    def apply_chat_template(self, messages, add_generation_prompt=True, tokenize=False):
        assert not tokenize and add_generation_prompt and len(messages) == 1 and messages[0]["role"] == "user"
        body = "".join("<|image|>" if c["type"] == "image" else c["text"] for c in messages[0]["content"])
        self.question = [c["text"] for c in messages[0]["content"] if c["type"] == "text"][0]
        return f"<|begin_of_text|><|start_header_id|>user<|end_header_id|>\n\n{body}<|eot_id|><|start_header_id|>assistant<|end_header_id|>\n\n"

    def __call__(self, texts):
        self.seen = texts[0]
        ids, pat = [], re.compile("|".join(re.escape(k) for k in sorted(self.SPECIAL, key=len, reverse=True)))
        pos = 0
        for m in pat.finditer(texts[0]):
            ids += [1000 + ord(ch) for ch in texts[0][pos:m.start()]]
            ids.append(self.SPECIAL[m.group(0)])
            pos = m.end()
        ids += [1000 + ord(ch) for ch in texts[0][pos:]]
        return {"input_ids": [ids]}


def _masks(n, h=300, w=400):
    out = []
    for i in range(n):
        m = np.zeros((h, w), bool)
        m[20 + 30 * i: 120 + 30 * i, 30 + 40 * i: 200 + 40 * i] = True
        out.append(m)
    return out


def test_single_region_request_matches_the_reference_prompt_and_layout():
    spec, tok = gar_tiny(), StubTokenizer()
    img = np.zeros((300, 400, 3), np.uint8)
    req = single_region_request(spec, tok, img, _masks(1)[0], prompt_token="<Prompt2>", max_num_tiles=4, dtype=torch.float32)
    # eval_dataset.py:108-113
    crop = "<|reserved_special_token_4|>"
    assert tok.question == ("There are some objects I am curious about: <Prompt2>;\n<Prompt2>: " + crop * 256 +
                            "Describe this masked region in detail.")
    ncw, nch = req["aspect_ratios"][0].tolist()
    T = 1 + ncw * nch
    assert tok.seen.count("<|image|>") == T * spec.tokens_per_tile          # processing_perception_lm.py:200-220
    ids = req["input_ids"][0]
    assert int((ids == spec.image_token_id).sum()) == T * 256 and int((ids == 128008).sum()) == 256
    assert list(req["bboxes"][0]) == ["128008"] and req["pixel_values"].shape[0] == T
    assert req["attention_mask"].shape == req["input_ids"].shape and bool((req["attention_mask"] == 1).all())
    raw = single_region_request(spec, tok, img, _masks(1)[0], prompt_token="<Prompt2>", max_num_tiles=4, raw=True)
    assert torch.equal(raw["input_ids"], req["input_ids"]) and raw["image_u8"].shape == (3, 300, 400)
    assert raw["idmap_u8"].dtype == torch.uint8 and set(raw["idmap_u8"].unique().tolist()) == {2, 5}


def test_multi_region_request_reproduces_the_reference_layout_and_quirks():
    spec, tok = gar_tiny(), StubTokenizer()
    case = GOLD["multi"][1]                         # 3 regions; "order" = the reference's set-iteration order there
    n, order = case["n"], case["order"]
    img = np.zeros((case["h"], case["w"], 3), np.uint8)
    from test_preprocess import _rect_masks

    masks = _rect_masks(case["seed"], case["h"], case["w"], n)
    q = "What is the relationship between " + " and ".join(f"<Prompt{i}>" for i in range(n)) + "?"
    req = multi_region_request(spec, tok, img, masks, q, max_num_tiles=4, dtype=torch.float32, prompt_order=order)
    # eval_dataset.py:207-221: object list, one "<PromptK>: <crop>*256\n" line per object in visiting order, the question
    want = "There are some objects I am curious about: " + "".join(f"<Prompt{k}>; " for k in order) + "\n"
    want += "".join(f"<Prompt{k}>: " + f"<|reserved_special_token_{k + 2}|>" * 256 + "\n" for k in order) + "\n" + q
    assert tok.question == want
    assert {k: list(map(float, v)) for k, v in req["bboxes"][0].items()} == case["bboxes"]     # stale last-mask boxes
    clean = multi_region_request(spec, tok, img, masks, q, max_num_tiles=4, dtype=torch.float32, reference_quirks=False)
    assert len(set(clean["bboxes"][0].values())) == n
    starts = [tok.question.index(f"<Prompt{k}>: ") for k in range(n)]
    assert starts == sorted(starts)                                                             # index order without quirks
