"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/idmap_goldens.json from the reference's own request builders.

Runs in the build container only (needs /root/reference).  `evaluation/eval_dataset.py` cannot be imported here
(`-1 * np.ones(..., dtype=np.uint8)` raises OverflowError under numpy 2; the reference pins numpy 1.26.4, where the
product promotes to int16), so `SingleRegionCaptionDataset._parse_annotations` (:56-99) and
`MultiRegionDataset._parse_annotations` (:192-271) are lifted *by source text* and executed unmodified against a
numpy facade whose only difference is that promotion (`ones(uint8)` comes back as int16, exactly what numpy 1.x makes of
`-1 * ones(uint8)`).  The tokenizer is a stub that knows the ids every caller relies on (SURVEY.md §8 'Special ids').

What is pinned: the painted id map (sha256 + histogram), the bbox dict (incl. quirk 6: every region of the multi-region
builder gets the LAST mask's box) and the order in which the multi-region builder visits its `set` of prompt strings.
"""
import ast
import hashlib
import json
import os
import re
import textwrap
import types

import numpy as np

REF = "/root/reference/evaluation/eval_dataset.py"
RESERVED = {2: 128004, 3: 128005, 4: 128008, 5: 128010, 6: 128011}
PROMPT_IDS = {"<Prompt0>": 128256, "<Prompt1>": 128257, "<Prompt2>": 128258, "<Prompt3>": 128259, "<Prompt4>": 128260,
              "<NO_Prompt>": 128261}


class _Tok:
    def convert_tokens_to_ids(self, tok):
        if tok in PROMPT_IDS:
            return PROMPT_IDS[tok]
        m = re.match(r"<\|reserved_special_token_(\d+)\|>", tok)
        return RESERVED[int(m.group(1))]


class _Proc:
    tokenizer = _Tok()


class _Img:
    def __init__(self, w, h):
        self.width, self.height, self.size = w, h, (w, h)

    def __array__(self, dtype=None, copy=None):
        return np.zeros((self.height, self.width, 3), dtype=np.uint8)


def numpy1_facade():
    f = types.SimpleNamespace(**{k: getattr(np, k) for k in dir(np) if not k.startswith("__")})

    def ones(shape, dtype=None):
        # numpy 1.x: (-1 * ones(uint8)) is int16; numpy 2 raises OverflowError instead
        return np.ones(shape, dtype=np.int16 if dtype == np.uint8 else dtype)

    f.ones = ones
    return f


def lift(cls_name):
    src = open(REF).read()
    tree = ast.parse(src)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls_name][0]
    fn = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "_parse_annotations"][0]
    code = textwrap.dedent(ast.get_source_segment(src, fn))
    ns = {"np": numpy1_facade(), "re": re,
          "Image": types.SimpleNamespace(fromarray=lambda a: a, NEAREST=0)}
    exec(code, ns)
    return ns["_parse_annotations"]


def rect_masks(seed, h, w, n):
    """the same seeded rectangles the test rebuilds"""
    g = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        hh, ww = int(g.integers(h // 8, h // 2)), int(g.integers(w // 8, w // 2))
        y0, x0 = int(g.integers(0, h - hh)), int(g.integers(0, w - ww))
        m = np.zeros((h, w), dtype=bool)
        m[y0:y0 + hh, x0:x0 + ww] = True
        out.append(m)
    return out


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a.astype(np.uint8)).tobytes()).hexdigest()


def main():
    single = lift("SingleRegionCaptionDataset")
    multi = lift("MultiRegionDataset")
    vp_ids = {k: v - 128256 for k, v in PROMPT_IDS.items()}
    out = {"single": [], "multi": []}
    for seed, (h, w), k in ((1, (300, 400), 0), (2, (517, 333), 3), (3, (64, 64), 4)):
        m = rect_masks(seed, h, w, 1)[0]
        me = types.SimpleNamespace(image=_Img(w, h), mask=m, prompt_token=f"<Prompt{k}>", visual_prompt_ids=vp_ids,
                                   processor=_Proc())
        d = single(me)
        fm = np.asarray(d["visual_prompt"])
        out["single"].append({"seed": seed, "h": h, "w": w, "prompt": k, "sha256": digest(fm),
                              "hist": {str(int(v)): int(c) for v, c in zip(*np.unique(fm, return_counts=True))},
                              "bboxes": {kk: [float(x) for x in vv] for kk, vv in d["bboxes"].items()}})
    for seed, (h, w), n in ((11, (300, 400), 2), (12, (480, 640), 3), (13, (200, 900), 4), (14, (128, 128), 5)):
        masks = rect_masks(seed, h, w, n)
        q = "What is the relationship between " + " and ".join(f"<Prompt{i}>" for i in range(n)) + "?"
        me = types.SimpleNamespace(image=_Img(w, h), masks=[mm.astype(np.uint8) for mm in masks], question_str=q,
                                   visual_prompt_ids=vp_ids, prompt_number=5, processor=_Proc())
        d = multi(me)
        fm = np.asarray(d["visual_prompt"])
        # visiting order of the reference's `set` of prompt strings, read back from the prompt it built
        order = [int(x) for x in re.findall(r"<Prompt(\d+)>; ", d["prompt"].split("\n")[0])]
        out["multi"].append({"seed": seed, "h": h, "w": w, "n": n, "order": order, "sha256": digest(fm),
                             "hist": {str(int(v)): int(c) for v, c in zip(*np.unique(fm, return_counts=True))},
                             "bboxes": {kk: [float(x) for x in vv] for kk, vv in d["bboxes"].items()}})
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "idmap_goldens.json")
    json.dump(out, open(dst, "w"), indent=1)
    print("wrote", dst, len(out["single"]), "single-region,", len(out["multi"]), "multi-region cases")


if __name__ == "__main__":
    main()
