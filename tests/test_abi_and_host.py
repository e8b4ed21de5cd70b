"""CPU-side checks: the C-ABI library builds, loads without a GPU and exports every symbol that
include/gar_b200.h declares; host-side request plumbing (config round trip, synthetic requests, box
arithmetic) matches the reference's conventions; the product refuses to run without CUDA."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from grasp_any_region_b200 import _lib
from grasp_any_region_b200.config import gar_1b, gar_8b, gar_tiny
from grasp_any_region_b200.configuration_gar import GARConfig
from grasp_any_region_b200.engine import roi_boxes_f32
from grasp_any_region_b200.synth import make_request

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from grasp_any_region_b200.build import build_library

    build_library()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "gar_b200.h")).read()
    declared = set(re.findall(r"\b(gar_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in gar_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.gar_abi_version() == 1


def test_ctypes_signatures_match_header_arity():
    """Every binding in _lib.SIGNATURES passes exactly as many arguments as the prototype in gar_b200.h declares
    (a mismatch would only show up as stack garbage on the GPU box)."""
    header = open(os.path.join(ROOT, "include", "gar_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    seen = {}
    for m in re.finditer(r"\b(gar_[a-z0-9_]+)\s*\(", header):
        depth, i = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(header[i], 0)
            i += 1
        if header[i:].lstrip()[:1] != ";":
            continue  # not a prototype (e.g. a mention inside a macro)
        params = header[m.end():i - 1].strip()
        seen[m.group(1)] = 0 if params in ("", "void") else params.count(",") + 1
    assert set(seen) == set(_lib.SIGNATURES)
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        assert len(argtypes) == seen[name], f"{name}: header declares {seen[name]} parameters, ctypes binds {len(argtypes)}"


def test_desc_struct_layout_matches_header():
    # 9 ints/floats, 8 ints, float, int, (pad), int64, 8 x int64, 4 ints
    assert ctypes.sizeof(_lib.GarModelDesc) == 4 * 20 + 8 + 64 + 24  # 5 capacity ints + tail padding
    assert _lib.GarModelDesc.image_token_id.offset == 80


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    from grasp_any_region_b200 import GARModel

    m = GARModel(GARConfig.from_spec(gar_tiny()))
    with pytest.raises(_lib.GarError):
        m.engine()
    d = _lib.GarModelDesc()
    d.vit_dim, d.vit_heads, d.img_size, d.patch_size, d.llm_head_dim = 128, 2, 448, 14, 64
    h = ctypes.c_void_p()
    assert lib.gar_engine_create(ctypes.byref(d), ctypes.byref(h)) != 0
    assert b"CUDA" in lib.gar_last_error()


def test_config_round_trip():
    for spec in (gar_1b(), gar_8b(), gar_tiny()):
        cfg = GARConfig.from_spec(spec)
        assert cfg.spec.to_dict() == spec.to_dict()
        again = GARConfig(**{k: v for k, v in cfg.to_dict().items()
                             if k in ("mllm_config", "prompt_numbers", "crop_tokens_ids")})
        assert again.spec.to_dict() == spec.to_dict()
    assert GARConfig.from_spec(gar_1b()).kernel_size == [14, 14]
    with pytest.raises(AssertionError):
        GARConfig(prompt_numbers=4)


@pytest.mark.parametrize("n_masks,mnt", [(1, 16), (4, 16), (2, 4)])
def test_synthetic_request_layout(n_masks, mnt):
    spec = gar_1b()
    r = make_request(spec, seed=3, n_masks=n_masks, max_num_tiles=mnt)
    T = int(np.floor(np.sqrt(mnt))) ** 2 + 1
    assert r["pixel_values"].shape == (T, 3, 448, 448) and r["global_mask_values"].shape == (T, 3, 448, 448)
    ids = r["input_ids"][0]
    assert int((ids == spec.image_token_id).sum()) == T * 256
    for k in range(n_masks):
        pos = (ids == spec.crop_tokens_ids[k]).nonzero().flatten()
        assert pos.numel() == 256 and int(pos.max() - pos.min()) == 255
        x1, y1, x2, y2 = r["bboxes"][0][str(spec.crop_tokens_ids[k])]
        assert 0 <= x1 <= x2 < 1 and 0 <= y1 <= y2 < 1
    # ids decode exactly in bf16 (reference quirk 7)
    dec = torch.clamp(torch.round((r["global_mask_values"] + 1.0) / 2.0 * 255.0).long(), 0, spec.prompt_numbers)
    assert set(dec.unique().tolist()) <= set(range(spec.prompt_numbers + 1))
    assert spec.prompt_numbers in dec.unique().tolist() and 0 in dec.unique().tolist()
    if mnt == 16 and n_masks == 1:
        assert ids.numel() == 4672


def test_roi_box_host_arithmetic_matches_oracle():
    from oracle.gar_oracle import roi_box_f32

    spec = gar_1b()
    boxes = {str(spec.crop_tokens_ids[0]): (0.1234567, 0.2, 0.7654321, 0.9),
             str(spec.crop_tokens_ids[3]): (0.0, 0.0, 0.999, 0.999)}
    rois, has, scale = roi_boxes_f32(spec, boxes, 4, 4)
    assert has == [1, 0, 0, 1, 0]
    for k in (0, 3):
        want, s = roi_box_f32(boxes[str(spec.crop_tokens_ids[k])], 64, 64)
        assert rois[k] == want.tolist() and s == scale


def test_checkpoint_round_trip_keeps_reference_key_layout(tmp_path):
    """N2: save_pretrained / from_pretrained with the reference's HF state-dict names
    (mllm.model.vision_tower.timm_model.*, ...language_model.*, mllm.lm_head, mask_patch_embedding)."""
    from grasp_any_region_b200 import GARModel
    from oracle.gar_oracle import build_oracle

    spec = gar_tiny()
    m = GARModel(GARConfig.from_spec(spec)).random_init_(3)
    assert set(m.state_dict()) == set(build_oracle(spec).state_dict())
    m.save_pretrained(tmp_path)
    m2 = GARModel.from_pretrained(tmp_path, torch_dtype=torch.bfloat16)
    a, b = m.state_dict(), m2.state_dict()
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert m2.config.prompt_numbers == 5 and m2.crop_tokens_ids == spec.crop_tokens_ids
    assert m2.lm_head.weight.shape == (spec.text.vocab_size, spec.text.hidden_size)


def test_gemm_raster_groups_cover_every_tile_once(lib):
    """Host-side view of the tcgen05 GEMM's tile walk (gar_gemm_raster_probe, no CUDA call): every (row block, column
    panel) is visited exactly once, n runs fastest inside a group of panels, a ragged last group is handled, and the
    L2 working-set rule picks the groups DESIGN.md quotes for the production shapes."""
    import ctypes as C

    def probe(M, N, K, tile_m, block_n, fly, tile=None):
        ng, m, n = C.c_int(), C.c_int(), C.c_int()
        rc = lib.gar_gemm_raster_probe(M, N, K, tile_m, block_n, fly, 0 if tile is None else tile, C.byref(ng),
                                       None if tile is None else C.byref(m), None if tile is None else C.byref(n))
        assert rc == 0
        return ng.value, m.value, n.value

    try:
        # production shapes, CTA-pair kernel (256 x 256 tiles, 74 pairs in flight)
        assert probe(37376, 16384, 2048, 256, 256, 74)[0] == 32    # GAR-1B gate/up: 64 panels of 1 MB -> 2 groups
        assert probe(37376, 2048, 8192, 256, 256, 74)[0] == 8      # GAR-1B down: no working set fits -> one group
        assert probe(139400, 4096, 1024, 256, 256, 74)[0] == 16    # ViT fc1: everything fits -> one group
        assert probe(139400, 1024, 4096, 256, 256, 74)[0] == 4     # ViT fc2: long K -> one group
        assert probe(9400, 28672, 4096, 256, 256, 74)[0] == 14     # GAR-8B gate/up: 112 panels of 2 MB -> 8 groups
        assert lib.gar_set_global_option(b"gemm_group_kb", 0) == 0
        assert probe(37376, 16384, 2048, 256, 256, 74)[0] == 64    # grouping off
        for M, N, K, tm, bn, fly, force in [(2300, 2816, 256, 256, 256, 74, 3), (1000, 3328, 512, 128, 128, 148, 4),
                                            (4097, 2048, 1024, 128, 64, 148, 5), (700, 512, 512, 128, 256, 148, 1)]:
            assert lib.gar_set_global_option(b"gemm_group_force", force) == 0
            num_m, num_n = -(-M // tm), -(-N // bn)
            seen, prev = set(), None
            for t in range(num_m * num_n):
                ng, m, n = probe(M, N, K, tm, bn, fly, t)
                assert ng == min(force, num_n) and 0 <= m < num_m and 0 <= n < num_n
                seen.add((m, n))
                if prev is not None and prev[1] // ng == n // ng and prev[0] == m:
                    assert n == prev[1] + 1  # n fastest inside a group
                prev = (m, n)
            assert len(seen) == num_m * num_n
    finally:
        lib.gar_set_global_option(b"gemm_group_force", 0)
        lib.gar_set_global_option(b"gemm_group_kb", 40960)

