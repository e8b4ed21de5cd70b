"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy float32 restatement of ``torchvision.ops.roi_align`` (aligned=True, one RoI, batch 1) as
the reference calls it for the feature replay (modeling_gar.py:389-396).  torchvision's C++
source is not in /root/reference (un-vendored dependency, version unpinned by the reference;
0.26.0 in this image) -- this follows its published algorithm (ops/cpu/roi_align_kernel.cpp /
ops/roi_align.py:_roi_align) and is pinned BIT-EXACT against the installed CPU operator in
tests/test_roi_replay_ref.py.  The CUDA replay kernel follows this file operation by operation
(no FMA contraction) so its rows are bit-identical after the bf16 cast.
"""
import numpy as np

f32 = np.float32


def roi_align_sample_table(roi, spatial_scale, H, W, pooled=16, sampling_ratio=2):
    """Per output bin (ph, pw) and sample (iy, ix): the 4 corner indices + 4 fp32 weights.
    Returns pos (pooled, pooled, s, s, 4) int32 [linear index y*W+x, or -1 = contributes 0] and
    w (pooled, pooled, s, s, 4) float32."""
    scale = f32(spatial_scale)
    off = f32(0.5)
    roi = np.asarray(roi, dtype=f32)
    start_w = f32(roi[1] * scale) - off
    start_h = f32(roi[2] * scale) - off
    end_w = f32(roi[3] * scale) - off
    end_h = f32(roi[4] * scale) - off
    roi_w = f32(end_w - start_w)
    roi_h = f32(end_h - start_h)
    bin_h = f32(roi_h / f32(pooled))
    bin_w = f32(roi_w / f32(pooled))
    s = sampling_ratio
    pos = np.full((pooled, pooled, s, s, 4), -1, dtype=np.int32)
    wts = np.zeros((pooled, pooled, s, s, 4), dtype=f32)
    for ph in range(pooled):
        for pw in range(pooled):
            for iy in range(s):
                yy = f32(f32(start_h + f32(f32(ph) * bin_h)) + f32(f32(f32(iy + 0.5) * bin_h) / f32(s)))
                for ix in range(s):
                    xx = f32(f32(start_w + f32(f32(pw) * bin_w)) + f32(f32(f32(ix + 0.5) * bin_w) / f32(s)))
                    y, x = yy, xx
                    if y < -1.0 or y > H or x < -1.0 or x > W:
                        continue
                    if y <= 0:
                        y = f32(0)
                    if x <= 0:
                        x = f32(0)
                    y_low, x_low = int(y), int(x)
                    if y_low >= H - 1:
                        y_high = y_low = H - 1
                        y = f32(y_low)
                    else:
                        y_high = y_low + 1
                    if x_low >= W - 1:
                        x_high = x_low = W - 1
                        x = f32(x_low)
                    else:
                        x_high = x_low + 1
                    ly = f32(y - f32(y_low))
                    lx = f32(x - f32(x_low))
                    hy = f32(f32(1.0) - ly)
                    hx = f32(f32(1.0) - lx)
                    pos[ph, pw, iy, ix] = (y_low * W + x_low, y_low * W + x_high, y_high * W + x_low,
                                           y_high * W + x_high)
                    wts[ph, pw, iy, ix] = (f32(hy * hx), f32(hy * lx), f32(ly * hx), f32(ly * lx))
    return pos, wts


def roi_align_ref(fmap, roi, spatial_scale, pooled=16, sampling_ratio=2):
    """fmap (C,H,W) float32 -> (C, pooled, pooled) float32, same accumulation order as the operator."""
    C, H, W = fmap.shape
    pos, wts = roi_align_sample_table(roi, spatial_scale, H, W, pooled, sampling_ratio)
    flat = fmap.reshape(C, H * W).astype(f32)
    out = np.zeros((C, pooled, pooled), dtype=f32)
    count = f32(max(sampling_ratio * sampling_ratio, 1))
    for ph in range(pooled):
        for pw in range(pooled):
            acc = np.zeros(C, dtype=f32)
            for iy in range(sampling_ratio):
                for ix in range(sampling_ratio):
                    p = pos[ph, pw, iy, ix]
                    if p[0] < 0:
                        continue
                    w = wts[ph, pw, iy, ix]
                    v = f32(w[0]) * flat[:, p[0]]
                    v = v + f32(w[1]) * flat[:, p[1]]
                    v = v + f32(w[2]) * flat[:, p[2]]
                    v = v + f32(w[3]) * flat[:, p[3]]
                    acc = acc + v
            out[:, ph, pw] = acc / count
    return out
