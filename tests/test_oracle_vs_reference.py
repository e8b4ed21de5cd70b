"""Pins the oracle restatement (oracle/gar_oracle.py) to the reference's OWN modules.

tests/golden/ref_tiny_*.npz were produced by oracle/ref_harness.py, which executes the
unmodified reference GARModel.generate + vendored PerceptionLM on CPU fp32 in the build
container.  Inputs are regenerated from the seed; weights from init seed 0.
"""
import os

import numpy as np
import pytest
import torch

from grasp_any_region_b200.config import gar_tiny
from grasp_any_region_b200.synth import make_request
from oracle.gar_oracle import build_oracle
from oracle.ref_harness import CASES, compact


@pytest.fixture(scope="module")
def tiny_oracle():
    torch.set_num_threads(os.cpu_count())
    return build_oracle(gar_tiny(), seed=0, dtype=torch.float32)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_run(name, tiny_oracle, golden_dir):
    spec = gar_tiny()
    n_masks, mnt, isz, ntok, seed = CASES[name]
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    req = make_request(spec, seed=seed, n_masks=n_masks, image_size=isz, max_num_tiles=mnt,
                       dtype=torch.float32)
    taps = tiny_oracle.generate(req["pixel_values"], req["global_mask_values"], req["aspect_ratios"],
                                req["bboxes"], req["input_ids"], max_new_tokens=ntok)
    mine = compact(taps, req, spec)
    assert np.array_equal(mine["embed_rows"], gold["embed_rows"])
    # integer / index outputs: exact
    assert np.array_equal(mine["sequences"], gold["sequences"])
    # fp32 taps: same math, different op grouping only -> tight tolerance
    for k, tol in (("mask_embeds", 1e-6), ("image_features", 2e-5), ("inputs_embeds", 2e-5),
                   ("logits_strided", 1e-4), ("logits_top_val", 1e-4)):
        d = np.abs(mine[k] - gold[k]).max()
        assert d <= tol, f"{k}: max abs diff {d}"
    assert np.array_equal(mine["logits_top_idx"][:8], gold["logits_top_idx"][:8])


def test_replay_rows_bit_exact(tiny_oracle, golden_dir):
    """The RoI replay rows (index + fp32 lerp, then cast) must be bit-identical (north_star)."""
    spec = gar_tiny()
    name = "ref_tiny_3mask"
    n_masks, mnt, isz, ntok, seed = CASES[name]
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    req = make_request(spec, seed=seed, n_masks=n_masks, image_size=isz, max_num_tiles=mnt,
                       dtype=torch.float32)
    taps = tiny_oracle.prefill_embeds(req["pixel_values"], req["global_mask_values"],
                                      req["aspect_ratios"], req["bboxes"], req["input_ids"])
    ids = req["input_ids"][0]
    rows = torch.from_numpy(gold["embed_rows"])
    is_crop = torch.isin(ids[rows], torch.tensor(spec.crop_tokens_ids))
    mine = taps["inputs_embeds"][0, rows][is_crop].numpy()
    ref = gold["inputs_embeds"][is_crop.numpy()]
    assert mine.shape[0] == 3 * 256
    # image_features feeding the replay agree to ~1e-6, so allow that much; the structure
    # (which cells, which weights) is exact
    assert np.abs(mine - ref).max() <= 2e-5


@pytest.mark.parametrize("name", ["ref_tiny_plm_replay", "ref_tiny_plm_video"])
def test_oracle_plm_replay_matches_reference_run(name, tiny_oracle, golden_dir):
    """the replay branches of the vendored PLM's prepare_inputs_for_generation
    (modeling_perception_lm.py:648-852), executed unmodified by oracle/ref_harness.py, against the restatement
    GAROracle.assemble_plm -- including the image branch's quirk that only the last crop token keeps its replay."""
    from oracle.ref_harness import PLM_CASES, compact_plm, frame_crop_token_id

    spec = gar_tiny()
    video, n_masks, mnt, seed = PLM_CASES[name]
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    req = make_request(spec, seed=seed, n_masks=n_masks, image_size=1024, max_num_tiles=mnt, dtype=torch.float32)
    with torch.no_grad():
        feats = tiny_oracle.image_features(req["pixel_values"], None)
        T = feats.shape[0]
        emb = tiny_oracle.assemble_plm(req["input_ids"], feats, req["aspect_ratios"], req["bboxes"],
                                       crop_tokens=spec.crop_tokens_ids, video=video,
                                       frame_crop_ids=[frame_crop_token_id(f) for f in range(T)])
    mine = compact_plm(dict(image_features=feats, inputs_embeds=emb), req, spec)
    assert np.array_equal(mine["embed_rows"], gold["embed_rows"])
    assert np.abs(mine["image_features"] - gold["image_features"]).max() <= 2e-5
    assert np.abs(mine["inputs_embeds"] - gold["inputs_embeds"]).max() <= 2e-5
    ids = req["input_ids"][0]
    rows = torch.from_numpy(gold["embed_rows"])
    table = tiny_oracle.mllm.model.language_model.embed_tokens.weight
    present = [c for c in spec.crop_tokens_ids if bool((ids == c).any())]
    assert len(present) == 3
    for c in present:
        sel = (ids[rows] == c).numpy()
        kept_embedding = np.abs(gold["inputs_embeds"][sel] - table[c].detach().numpy()[None]).max() <= 1e-6
        # image branch: only the LAST crop token present is replayed, the others still hold embed_tokens(crop id);
        # video branch: every frame's crop run is replayed
        assert kept_embedding == ((not video) and c != present[-1]), (c, kept_embedding)
