"""GPU parity of every per-kernel C-ABI entry point against plain torch restatements of the op the
reference executes (same bf16 rounding points).  Calls go through libgar_b200.so (ctypes)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():  # collected on the CPU box, executed only with -m gpu on a B200
    pytest.skip("needs a CUDA device", allow_module_level=True)

from grasp_any_region_b200 import ops  # noqa: E402

DEV = "cuda:0"
bf = torch.bfloat16


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(bf).to(DEV)


def r(x):  # one bf16 rounding point
    return x.to(bf).float()


def gemm_ref(a, b, epi, bias=None, gamma=None, res=None):
    acc = a.float() @ b.float().t()
    if epi in (ops.EPI_BIAS, ops.EPI_BIAS_GELU, ops.EPI_BIAS_LS_RES):
        acc = acc + bias.float()
    y = r(acc)
    if epi == ops.EPI_BIAS_GELU:
        y = r(F.gelu(y))
    if epi == ops.EPI_BIAS_LS_RES:
        y = r(res.float() + r(gamma.float() * y))
    if epi == ops.EPI_RES:
        y = r(res.float() + y)
    return y


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 512, 256), (1, 256, 128), (8, 1024, 512), (2050, 768, 1024),
                                   (129, 96, 640), (4100, 3072, 1024)])
@pytest.mark.parametrize("block_n", [0, 64, 128, 256])
def test_gemm_plain(M, N, K, block_n):
    a, b = rnd(M, K, seed=1), rnd(N, K, scale=0.05, seed=2)
    out = ops.gemm(a, b, block_n=block_n)
    torch.cuda.synchronize()
    ref = gemm_ref(a, b, ops.EPI_NONE)
    err = (out.float() - ref).abs().max().item()
    tol = 2e-2 * max(1.0, ref.abs().max().item())  # <= 1 bf16 ulp flips from accumulation order
    assert err <= tol, (err, tol)
    mism = (out.float() != ref).float().mean().item()
    assert mism < 0.02, f"{mism:.4f} of outputs differ from the fp32-accumulated reference by a bf16 ulp"


@pytest.mark.parametrize("M,N,K,epi", [(4100, 1024, 4096, ops.EPI_BIAS_LS_RES),   # ViT fc2 (+ LayerScale residual)
                                       (4672, 2048, 8192, ops.EPI_RES),           # Llama-3.2-1B down_proj (+ residual)
                                       (2300, 4096, 14336, ops.EPI_RES),          # Llama-3.1-8B down_proj
                                       (3000, 1536, 8960, ops.EPI_BIAS_LS_RES)])  # PE-G fc2
def test_gemm_long_k_production_shapes(M, N, K, epi):
    """The big-M tcgen05 GEMM at the longest reduction depths of the path (64..224 k-blocks through the TMA ring and
    one TMEM accumulator): fp32-accumulated reference on the same bf16 inputs."""
    a, b = rnd(M, K, scale=0.5, seed=11), rnd(N, K, scale=0.02, seed=12)
    bias, gamma, res = rnd(N, scale=0.1, seed=13), rnd(N, scale=0.2, seed=14), rnd(M, N, seed=15)
    out = ops.gemm(a, b, epi=epi, bias=bias, gamma=gamma, res=res)
    torch.cuda.synchronize()
    ref = gemm_ref(a, b, epi, bias, gamma, res)
    err = (out.float() - ref).abs().max().item()
    assert err <= 4e-2 * max(1.0, ref.abs().max().item()), err
    assert (out.float() != ref).float().mean().item() < 0.03


@pytest.mark.parametrize("epi", [ops.EPI_BIAS, ops.EPI_BIAS_GELU, ops.EPI_BIAS_LS_RES, ops.EPI_RES])
@pytest.mark.parametrize("M,N,K", [(333, 512, 256), (1025, 1024, 1024), (4000, 2048, 512)])
def test_gemm_epilogues(epi, M, N, K):
    a, b = rnd(M, K, seed=3), rnd(N, K, scale=0.05, seed=4)
    bias, gamma, res = rnd(N, scale=0.1, seed=5), rnd(N, scale=0.2, seed=6), rnd(M, N, seed=7)
    out = ops.gemm(a, b, epi=epi, bias=bias, gamma=gamma, res=res)
    torch.cuda.synchronize()
    ref = gemm_ref(a, b, epi, bias, gamma, res)
    err = (out.float() - ref).abs().max().item()
    assert err <= 4e-2 * max(1.0, ref.abs().max().item()), err
    assert (out.float() != ref).float().mean().item() < 0.02


@pytest.mark.parametrize("M,I,K", [(300, 512, 256), (1000, 8192, 512), (130, 192, 128)])
def test_gemm_fused_swiglu(M, I, K):
    a = rnd(M, K, seed=60)
    gw, uw = rnd(I, K, scale=0.05, seed=61), rnd(I, K, scale=0.05, seed=62)
    want = ops.swiglu(torch.cat([ops.gemm(a, gw), ops.gemm(a, uw)], 1).contiguous())
    got = ops.gemm(a, ops.interleave_gate_up(gw, uw), epi=ops.EPI_SWIGLU)
    torch.cuda.synchronize()
    assert got.shape == (M, I)
    assert torch.equal(got, want)


@pytest.mark.parametrize("epi", [ops.EPI_NONE, ops.EPI_BIAS_GELU, ops.EPI_BIAS_LS_RES, ops.EPI_SWIGLU])
@pytest.mark.parametrize("M,N,K", [(4097, 2048, 1024), (17425, 1024, 512), (2300, 4096, 256)])
def test_gemm_cta_pair_matches_single_cta(epi, M, N, K):
    """The cta_group::2 kernel (256x256 tiles over two SMs) accumulates every output in the same order as the
    single-CTA kernel, so the two must agree bit for bit."""
    from grasp_any_region_b200 import _lib

    a, b = rnd(M, K, seed=80), rnd(N, K, scale=0.05, seed=81)
    bias, gamma, res = rnd(N, scale=0.1, seed=82), rnd(N, scale=0.2, seed=83), rnd(M, N, seed=84)
    lib = _lib.load()
    _lib.check(lib.gar_set_global_option(b"gemm_2cta", 0))
    one = ops.gemm(a, b, epi=epi, bias=bias, gamma=gamma, res=res)
    _lib.check(lib.gar_set_global_option(b"gemm_2cta", 1))
    two = ops.gemm(a, b, epi=epi, bias=bias, gamma=gamma, res=res)
    torch.cuda.synchronize()
    assert torch.equal(one, two)


@pytest.mark.parametrize("epi", [ops.EPI_NONE, ops.EPI_BIAS_LS_RES, ops.EPI_SWIGLU])
@pytest.mark.parametrize("cta2", [0, 1])
@pytest.mark.parametrize("M,N,K,panels", [(4097, 2048, 1024, 2),    # 8 (or 16 / 32 narrower) panels in groups of 2
                                          (2300, 2816, 256, 3),     # 11 panels -> ragged last group
                                          (17425, 1024, 512, 1),    # one panel per group
                                          (1000, 3328, 512, 4)])    # 13 / 26 panels, ragged
def test_gemm_grouped_raster_matches_plain_order(epi, cta2, M, N, K, panels):
    """The raster only changes WHICH tile a CTA takes next (column panels in L2-sized groups): every output keeps its
    accumulation order, so results are bit-identical to the single-group (n-fastest) walk -- ragged last groups, 64 /
    128 / 256-wide tiles and the CTA-pair kernel included."""
    from grasp_any_region_b200 import _lib

    a, b = rnd(M, K, seed=90), rnd(N, K, scale=0.05, seed=91)
    bias, gamma, res = rnd(N, scale=0.1, seed=92), rnd(N, scale=0.2, seed=93), rnd(M, N, seed=94)
    lib = _lib.load()
    _lib.check(lib.gar_set_global_option(b"gemm_2cta", cta2))
    try:
        _lib.check(lib.gar_set_global_option(b"gemm_group_kb", 0))
        plain = ops.gemm(a, b, epi=epi, bias=bias, gamma=gamma, res=res)
        _lib.check(lib.gar_set_global_option(b"gemm_group_force", panels))
        grouped = ops.gemm(a, b, epi=epi, bias=bias, gamma=gamma, res=res)
        torch.cuda.synchronize()
    finally:
        _lib.check(lib.gar_set_global_option(b"gemm_group_force", 0))
        _lib.check(lib.gar_set_global_option(b"gemm_group_kb", 40960))
        _lib.check(lib.gar_set_global_option(b"gemm_2cta", 1))
    assert torch.equal(plain, grouped)


def test_gemm_inplace_residual():
    M, N, K = 700, 512, 512
    a, b, x = rnd(M, K, seed=8), rnd(N, K, scale=0.05, seed=9), rnd(M, N, seed=10)
    ref = gemm_ref(a, b, ops.EPI_RES, res=x)
    out = ops.gemm(a, b, epi=ops.EPI_RES, res=x, out=x)
    torch.cuda.synchronize()
    assert out.data_ptr() == x.data_ptr()
    assert (out.float() - ref).abs().max().item() <= 4e-2 * ref.abs().max().item()


@pytest.mark.parametrize("M,N,K", [(1, 256, 256), (8, 3072, 2048), (8, 2048, 8192), (16, 1024, 512), (3, 128512, 512),
                                   (8, 16384, 2048), (5, 384, 128)])
@pytest.mark.parametrize("with_res", [False, True])
def test_gemm_skinny(M, N, K, with_res):
    x, w = rnd(M, K, seed=40), rnd(N, K, scale=0.05, seed=41)
    res = rnd(M, N, seed=42) if with_res else None
    out = ops.gemm_skinny(x, w, epi=ops.EPI_RES if with_res else ops.EPI_NONE, res=res)
    torch.cuda.synchronize()
    ref = gemm_ref(x, w, ops.EPI_RES if with_res else ops.EPI_NONE, res=res)
    err = (out.float() - ref).abs().max().item()
    assert err <= 4e-2 * max(1.0, ref.abs().max().item()), err
    assert (out.float() != ref).float().mean().item() < 0.02
    # deterministic split-K: bitwise repeatable
    out2 = ops.gemm_skinny(x, w, epi=ops.EPI_RES if with_res else ops.EPI_NONE, res=res)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)


@pytest.mark.parametrize("M", [1, 5, 8])
def test_gemv_decode_chain(M):
    """o-proj(+res, sumsq) -> gate-up(norm + swiglu) -> down(+res, sumsq) -> head(norm): the decode layer tail."""
    D, I, V = 512, 1536, 1024
    att, h = rnd(M, D, seed=70), rnd(M, D, seed=71)
    wo, wg, wu = rnd(D, D, scale=0.05, seed=72), rnd(I, D, scale=0.05, seed=73), rnd(I, D, scale=0.05, seed=74)
    wd, wl = rnd(D, I, scale=0.05, seed=75), rnd(V, D, scale=0.05, seed=76)
    ln2 = (1 + rnd(D, scale=0.1, seed=77).float()).to(bf)
    lnf = (1 + rnd(D, scale=0.1, seed=78).float()).to(bf)

    def rms(x, w):
        x32 = x.float()
        return w * (x32 * torch.rsqrt(x32.pow(2).mean(-1, keepdim=True) + 1e-5)).to(bf)

    # reference chain built from the already-verified kernels' semantics in torch
    h1 = gemm_ref(att, wo, ops.EPI_RES, res=h).to(bf)
    n2 = rms(h1, ln2)
    act = (F.silu(r(n2.float() @ wg.float().t()).to(bf)) * r(n2.float() @ wu.float().t()).to(bf))
    h2 = gemm_ref(act, wd, ops.EPI_RES, res=h1).to(bf)
    logits = r(rms(h2, lnf).float() @ wl.float().t())

    gemv = ops.gemv_decode
    g1, sq1 = gemv(att, wo, epi=1, res=h, want_sumsq=True)
    assert (g1.float() - h1.float()).abs().max().item() <= 4e-2
    assert torch.allclose(sq1.sum(0)[:M], g1.float().pow(2).sum(-1), rtol=1e-4)
    gact, _ = gemv(g1, ops.interleave_gate_up(wg, wu), epi=2, norm_w=ln2, sumsq_in=sq1)
    n2g = rms(g1, ln2)
    actg = (F.silu(r(n2g.float() @ wg.float().t()).to(bf)) * r(n2g.float() @ wu.float().t()).to(bf))
    assert (gact.float() - actg.float()).abs().max().item() <= 3e-2 * max(1.0, actg.float().abs().max().item())
    assert (gact != actg).float().mean().item() < 0.03
    g2, sq2 = gemv(gact, wd, epi=1, res=g1, want_sumsq=True)
    gl, _ = gemv(g2, wl, epi=0, norm_w=lnf, sumsq_in=sq2)
    torch.cuda.synchronize()
    ref_l = r(rms(g2, lnf).float() @ wl.float().t())
    assert (gl.float() - ref_l).abs().max().item() <= 4e-2 * max(1.0, ref_l.abs().max().item())
    assert (gl.float() - logits).abs().max().item() <= 0.15 * max(1.0, logits.abs().max().item())
    # bitwise repeatable
    gl2, _ = gemv(g2, wl, epi=0, norm_w=lnf, sumsq_in=sq2)
    assert torch.equal(gl, gl2)


@pytest.mark.parametrize("N,K,epi", [(3072, 2048, 0), (2048, 8192, 1), (16384, 2048, 2), (4096, 14336, 1), (128512, 2048, 0)])
def test_gemv_decode_production_shapes(N, K, epi):
    """GAR-1B / GAR-8B decode shapes against fp32 torch."""
    M = 8
    x, w = rnd(M, K, seed=90), rnd(N, K, scale=0.02, seed=91)
    res = rnd(M, N, seed=92) if epi == 1 else None
    nw = (1 + rnd(K, scale=0.1, seed=93).float()).to(bf) if epi != 1 else None
    sq = None
    xin = x
    if nw is not None:
        sq = torch.zeros((K // 16, 8), dtype=torch.float32, device=DEV)
        sq[0, :M] = x.float().pow(2).sum(-1)
        x32 = x.float()
        xin = nw * (x32 * torch.rsqrt(x32.pow(2).mean(-1, keepdim=True) + 1e-5)).to(bf)
    out, osq = ops.gemv_decode(x, w, epi=epi, res=res, norm_w=nw, sumsq_in=sq, want_sumsq=epi == 1)
    acc = xin.float() @ w.float().t()
    if epi == 2:
        a3 = acc.reshape(M, N // 64, 2, 32)
        ref = (F.silu(r(a3[:, :, 0]).to(bf).float()).to(bf).float() * r(a3[:, :, 1])).reshape(M, N // 2)
    elif epi == 1:
        ref = r(acc) + res.float()
    else:
        ref = acc
    scale = max(1.0, ref.abs().max().item())
    assert (out.float() - ref).abs().max().item() <= 2e-2 * scale
    if epi == 1:
        assert torch.allclose(osq.sum(0)[:M], out.float().pow(2).sum(-1), rtol=1e-4)


def test_layernorm_rmsnorm():
    x = rnd(1000, 1024, seed=11)
    w, b = (1 + rnd(1024, scale=0.1, seed=12).float()).to(bf), rnd(1024, scale=0.1, seed=13)
    y = ops.layernorm(x, w, b, 1e-5)
    ref = F.layer_norm(x.float(), (1024,), w.float(), b.float(), 1e-5).to(bf)
    assert (y.float() - ref.float()).abs().max().item() <= 2e-2
    assert (y != ref).float().mean().item() < 0.01
    x2 = rnd(333, 2048, seed=14)
    w2 = (1 + rnd(2048, scale=0.1, seed=15).float()).to(bf)
    y2 = ops.rmsnorm(x2, w2, 1e-5)
    x32 = x2.float()
    ref2 = w2 * (x32 * torch.rsqrt(x32.pow(2).mean(-1, keepdim=True) + 1e-5)).to(bf)
    assert (y2.float() - ref2.float()).abs().max().item() <= 2e-2
    assert (y2 != ref2).float().mean().item() < 0.01
    # widths without a fixed-width kernel, the ViT-G width, and fewer rows than one warp sweep
    for rows, d in ((3, 512), (1500, 1536), (70, 4096), (5, 2048)):
        xa = rnd(rows, d, seed=16 + d)
        wa, ba = (1 + rnd(d, scale=0.1, seed=17).float()).to(bf), rnd(d, scale=0.1, seed=18)
        ya = ops.layernorm(xa, wa, ba, 1e-5)
        ra = F.layer_norm(xa.float(), (d,), wa.float(), ba.float(), 1e-5).to(bf)
        assert (ya.float() - ra.float()).abs().max().item() <= 3e-2, (rows, d)
        yb = ops.rmsnorm(xa, wa, 1e-5)
        x32 = xa.float()
        rb = wa * (x32 * torch.rsqrt(x32.pow(2).mean(-1, keepdim=True) + 1e-5)).to(bf)
        assert (yb.float() - rb.float()).abs().max().item() <= 3e-2, (rows, d)


def test_patchify_and_mask_decode():
    T, P = 3, 5
    px = rnd(T, 3, 448, 448, seed=16)
    ids = torch.randint(0, P + 2, (T, 1, 448, 448), generator=torch.Generator().manual_seed(1)).float()
    mv = (2.0 * ids / 255.0 - 1.0).expand(T, 3, 448, 448).contiguous().to(bf).to(DEV)
    a_img, a_mask = ops.patchify(px, mv, 14, 640, P)
    torch.cuda.synchronize()
    # reference: unfold == conv im2col with (c, ky, kx) column order
    cols = F.unfold(px.float(), 14, stride=14).transpose(1, 2).reshape(T * 1024, 588)
    assert torch.equal(a_img[:, :588].float(), cols)
    assert a_img[:, 588:].abs().max().item() == 0
    dec = torch.clamp(torch.round((mv + 1.0) / 2.0 * 255.0).long(), 0, P)
    binm = (dec != P).to(bf)
    mcols = F.unfold(binm.float(), 14, stride=14).transpose(1, 2).reshape(T * 1024, 588)
    assert torch.equal(a_mask[:, :588].float(), mcols)


def test_vit_embed_ln():
    T, ntok, d = 2, 1025, 256
    patch = rnd(T * 1024, d, seed=17)
    cls, pos = rnd(d, seed=18), rnd(ntok, d, scale=0.1, seed=19)
    w, b = (1 + rnd(d, scale=0.1, seed=20).float()).to(bf), rnd(d, scale=0.1, seed=21)
    x = ops.vit_embed_ln(patch, cls, pos, w, b, T, ntok, 1, 1e-5)
    full = torch.cat([cls.view(1, 1, d).expand(T, 1, d), patch.view(T, 1024, d)], 1)
    full = full + pos  # bf16 add
    ref = F.layer_norm(full.float(), (d,), w.float(), b.float(), 1e-5).to(bf).view(T * ntok, d)
    assert (x.float() - ref.float()).abs().max().item() <= 2e-2
    assert (x != ref).float().mean().item() < 0.01


def test_vit_rope():
    from grasp_any_region_b200.config import VisionSpec
    from grasp_any_region_b200.rope import vit_rope_table

    v = VisionSpec(embed_dim=128, depth=1, num_heads=2, mlp_dim=256)
    T, ntok, H, hd = 2, 1025, 2, 64
    qkv = rnd(T * ntok, 3 * H * hd, seed=22)
    table = vit_rope_table(v).to(bf).to(DEV)
    ref = qkv.clone().view(T, ntok, 3, H, hd)
    sin, cos = table[:, :hd], table[:, hd:]

    def rot(x):
        return torch.stack([-x[..., 1::2], x[..., ::2]], -1).reshape(x.shape)

    for part in (0, 1):
        x = ref[:, 1:, part]  # (T, 1024, H, hd)
        ref[:, 1:, part] = x * cos[None, :, None, :] + rot(x) * sin[None, :, None, :]
    out = ops.vit_rope_(qkv.clone(), table, ntok, 1, H, hd)
    torch.cuda.synchronize()
    assert torch.equal(out.view(T, ntok, 3, H, hd), ref)


def test_gemm_fused_rope2d_equals_gemm_then_rope():
    from grasp_any_region_b200.config import VisionSpec
    from grasp_any_region_b200.rope import vit_rope_table

    v = VisionSpec(embed_dim=256, depth=1, num_heads=4, mlp_dim=256)
    T, ntok, H, hd, d = 2, 1025, 4, 64, 256
    a, w, bias = rnd(T * ntok, d, seed=50), rnd(3 * d, d, scale=0.05, seed=51), rnd(3 * d, scale=0.1, seed=52)
    table = vit_rope_table(v).to(bf).to(DEV)
    want = ops.vit_rope_(ops.gemm(a, w, epi=ops.EPI_BIAS, bias=bias), table, ntok, 1, H, hd)
    got = ops.gemm_rope2d(a, w, bias, table, ntok, 1, hd, 2 * d)
    torch.cuda.synchronize()
    assert torch.equal(got, want)


def test_llama_rope_kv_and_swiglu_pool_argmax():
    from grasp_any_region_b200.config import TextSpec
    from grasp_any_region_b200.rope import llama_cos_sin

    t = TextSpec()
    hq, hkv, hd, S, max_seq = 8, 2, 64, 200, 256
    cos, sin = llama_cos_sin(t, max_seq)
    cos, sin = cos.to(bf).to(DEV), sin.to(bf).to(DEV)
    qkv = rnd(S, (hq + 2 * hkv) * hd, seed=23)
    kc = torch.zeros(3, hkv, max_seq, hd, dtype=bf, device=DEV)
    vc = torch.zeros_like(kc)
    out = ops.llama_rope_kv_(qkv.clone(), hq, hkv, hd, cos, sin, kc, vc, max_seq, pos0=5, slot0=1)
    torch.cuda.synchronize()
    c = torch.cat([cos, cos], -1)[5:5 + S, None, :]
    s = torch.cat([sin, sin], -1)[5:5 + S, None, :]

    def rh(x):
        return torch.cat([-x[..., hd // 2:], x[..., :hd // 2]], -1)

    q = qkv[:, :hq * hd].view(S, hq, hd)
    k = qkv[:, hq * hd:(hq + hkv) * hd].view(S, hkv, hd)
    vv = qkv[:, (hq + hkv) * hd:].view(S, hkv, hd)
    assert torch.equal(out[:, :hq * hd].view(S, hq, hd), q * c + rh(q) * s)
    assert torch.equal(kc[1, :, 5:5 + S], (k * c + rh(k) * s).transpose(0, 1))
    assert torch.equal(vc[1, :, 5:5 + S], vv.transpose(0, 1))
    assert kc[0].abs().max().item() == 0 and kc[2].abs().max().item() == 0

    gu = rnd(77, 2 * 512, seed=24)
    act = ops.swiglu(gu)
    ref = F.silu(gu[:, :512]) * gu[:, 512:]
    assert torch.equal(act, ref)
    gi = torch.stack([gu[:, :512].reshape(77, 16, 32), gu[:, 512:].reshape(77, 16, 32)], 2).reshape(77, 1024)
    assert torch.equal(ops.swiglu(gi.contiguous(), interleaved=True), ref)

    x = rnd(2 * 1025, 256, seed=25)
    y = ops.pool2x2(x, 2, 1025, 1, 32)
    xr = x.view(2, 1025, 256)[:, 1:].permute(0, 2, 1).reshape(2, 256, 32, 32)
    refp = F.adaptive_avg_pool2d(xr, (16, 16)).flatten(2).transpose(1, 2)
    assert (y.float() - refp.float()).abs().max().item() <= 1e-2
    assert (y != refp).float().mean().item() < 0.002

    lg = rnd(4, 1000, seed=26)
    lg[1, 17] = lg[1].max() + 1
    lg[2, 5] = lg[2, 900] = lg[2].max() + 2  # tie -> lowest index
    nxt = ops.argmax_step(lg, 999)
    torch.cuda.synchronize()
    assert nxt.tolist() == lg[:, :999].float().argmax(-1).tolist()
    assert nxt[2].item() == 5


# (keys per block, persistent): the BKV 64 configuration exists only as the per-tile kernel
@pytest.fixture(params=[(64, 0), (128, 0), (128, 1)], ids=["bkv64", "bkv128-per-tile", "bkv128-persistent"])
def attn_bkv(request):
    from grasp_any_region_b200 import _lib

    bkv, persist = request.param
    _lib.check(_lib.load().gar_set_global_option(b"attn_bkv", bkv))
    _lib.check(_lib.load().gar_set_global_option(b"attn_persist", persist))
    yield request.param
    _lib.check(_lib.load().gar_set_global_option(b"attn_bkv", 128))
    _lib.check(_lib.load().gar_set_global_option(b"attn_persist", 1))


@pytest.mark.parametrize("hd", [64, 128])
# ragged lengths: 1025 = 8*128 + 1 (PE ViT with cls token), one-row / one-key tails, lengths past 2048
@pytest.mark.parametrize("T,ntok,heads", [(1, 128, 1), (2, 1025, 2), (1, 1024, 4), (3, 300, 2), (1, 65, 1), (2, 160, 2),
                                          (1, 289, 1), (2, 258, 2), (1, 131, 3), (1, 132, 1), (1, 133, 1), (1, 2049, 1), (2, 129, 2)])
def test_attention_vit_layout(T, ntok, heads, hd, attn_bkv):
    qkv = rnd(T * ntok, 3 * heads * hd, seed=27)
    out = ops.attn_fwd_packed_qkv(qkv, T, ntok, heads, hd)
    torch.cuda.synchronize()
    q, k, v = qkv.view(T, ntok, 3, heads, hd).permute(2, 0, 3, 1, 4).float()
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(T * ntok, heads * hd)
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-2, err


@pytest.mark.parametrize("hd", [64, 128])
@pytest.mark.parametrize("S", [64, 128, 333, 1000])
def test_attention_causal_gqa_cache_layout(S, hd, attn_bkv):
    hq, hkv, max_seq = 8, 2, 1024
    q = rnd(S, (hq + 2 * hkv) * hd, seed=28)
    kc = torch.zeros(hkv, max_seq, hd, dtype=bf, device=DEV)
    vc = torch.zeros_like(kc)
    kc[:, :S] = rnd(hkv, S, hd, seed=29)
    vc[:, :S] = rnd(hkv, S, hd, seed=30)
    out = ops.attn_fwd_cache(q, q.stride(0), kc, vc, hq, hkv, hd, S, max_seq)
    torch.cuda.synchronize()
    qq = q[:, :hq * hd].view(S, hq, hd).transpose(0, 1).float()[None]
    kk = kc[:, :S].float().repeat_interleave(hq // hkv, 0)[None]
    vv = vc[:, :S].float().repeat_interleave(hq // hkv, 0)[None]
    ref = F.scaled_dot_product_attention(qq, kk, vv, is_causal=True)[0].transpose(0, 1).reshape(S, hq * hd)
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-2, err


@pytest.mark.parametrize("hd", [64, 128])
def test_attention_persistent_is_bit_identical_to_per_tile(hd):
    """The persistent kernel walks the same tiles with the same arithmetic: outputs must be equal bit for bit, on the
    ViT shape (many more tiles than resident CTAs, 1-row query tail, 1-key KV tail), on a batched causal GQA prefill
    (tiles of 1..N blocks, heaviest first) and back to back on one stream (the tile counters reset themselves)."""
    from grasp_any_region_b200 import _lib

    lib = _lib.load()
    T, ntok, heads = (40, 1025, 16) if hd == 64 else (12, 1025, 8)
    qkv = rnd(T * ntok, 3 * heads * hd, seed=41)
    S, hq, hkv, max_seq, nreq = 3000, 16, 4, 3072, 2  # 24 query blocks x 16 heads = 384 tiles per call > resident CTAs
    q = rnd(nreq * S, (hq + 2 * hkv) * hd, seed=42)
    kc, vc = rnd(nreq * hkv, max_seq, hd, seed=43), rnd(nreq * hkv, max_seq, hd, seed=44)
    outs = {}
    for persist in (0, 1):
        _lib.check(lib.gar_set_global_option(b"attn_persist", persist))
        a = [ops.attn_fwd_packed_qkv(qkv, T, ntok, heads, hd) for _ in range(3)]
        c = [ops.attn_fwd_cache(q[i * S:(i + 1) * S], q.stride(0), kc[i * hkv:(i + 1) * hkv], vc[i * hkv:(i + 1) * hkv],
                                hq, hkv, hd, S, max_seq) for i in range(nreq)]
        torch.cuda.synchronize()
        assert all(torch.equal(a[0], x) for x in a[1:])
        outs[persist] = (a[0], torch.cat(c))
    _lib.check(lib.gar_set_global_option(b"attn_persist", 1))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])
    qf, kf, vf = qkv.view(T, ntok, 3, heads, hd).permute(2, 0, 3, 1, 4)[:, :2].float()
    ref = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(2 * ntok, heads * hd)
    assert (outs[1][0][:2 * ntok].float() - ref).abs().max().item() <= 2e-2


@pytest.mark.parametrize("mma", [1, 0])
def test_decode_attention(mma):
    from grasp_any_region_b200 import _lib

    _lib.check(_lib.load().gar_set_global_option(b"decode_attn_mma", mma))
    B, hq, hkv, hd, max_seq = 3, 8, 2, 64, 700
    qkv = rnd(B, (hq + 2 * hkv) * hd, seed=31)
    kc, vc = rnd(B, hkv, max_seq, hd, seed=32), rnd(B, hkv, max_seq, hd, seed=33)
    pos = torch.tensor([0, 333, 698], dtype=torch.int32, device=DEV)
    out = ops.decode_attn(qkv, kc, vc, pos, hq, hkv, hd, max_seq)
    torch.cuda.synchronize()
    for b in range(B):
        n = int(pos[b]) + 1
        q = qkv[b, :hq * hd].view(hq, 1, hd).float()[None]
        k = kc[b, :, :n].float().repeat_interleave(hq // hkv, 0)[None]
        v = vc[b, :, :n].float().repeat_interleave(hq // hkv, 0)[None]
        ref = F.scaled_dot_product_attention(q, k, v)[0, :, 0].reshape(hq * hd)
        err = (out[b].float() - ref).abs().max().item()
        assert err <= 2e-2, (b, err)
    _lib.check(_lib.load().gar_set_global_option(b"decode_attn_mma", 1))


def test_decode_attention_fused_rope_matches_unfused():
    """RoPE + KV append fused into the decode attention kernel == llama_rope_kv followed by decode attention,
    bit for bit, including the cache rows it appends."""
    from grasp_any_region_b200.config import TextSpec
    from grasp_any_region_b200.rope import llama_cos_sin

    B, hq, hkv, hd, max_seq = 4, 8, 2, 64, 600
    cos, sin = llama_cos_sin(TextSpec(), max_seq)
    cos, sin = cos.to(bf).to(DEV), sin.to(bf).to(DEV)
    qkv = rnd(B, (hq + 2 * hkv) * hd, seed=90)
    kc0, vc0 = rnd(B, hkv, max_seq, hd, seed=91), rnd(B, hkv, max_seq, hd, seed=92)
    pos = torch.tensor([0, 17, 333, 598], dtype=torch.int32, device=DEV)
    for splits in (1, 5, 8):
        kc1, vc1, q1 = kc0.clone(), vc0.clone(), qkv.clone()
        ops.llama_rope_kv_(q1, hq, hkv, hd, cos, sin, kc1, vc1, max_seq, pos_dev=pos, slot0=-1)
        want = ops.decode_attn(q1, kc1, vc1, pos, hq, hkv, hd, max_seq, splits=splits)
        kc2, vc2 = kc0.clone(), vc0.clone()
        got = ops.decode_attn_rope(qkv.clone(), kc2, vc2, pos, hq, hkv, hd, max_seq, cos, sin, splits=splits)
        torch.cuda.synchronize()
        assert torch.equal(kc1, kc2) and torch.equal(vc1, vc2)
        assert torch.equal(got, want)


@pytest.mark.parametrize("hd", [64, 128])
def test_decode_attention_long_context_and_poisoned_tail(hd):
    """Fused RoPE + append + decode attention at the bench's context (~4.7k keys, GAR-1B / GAR-8B head layout, batch 8) against
    fp32 SDPA over the rotated cache; cache rows PAST every live length are poisoned with NaN (never-written memory)
    and must not reach the output; the appended rows must be bit-identical to the unfused RoPE + append kernel."""
    from grasp_any_region_b200.config import TextSpec
    from grasp_any_region_b200.rope import llama_cos_sin

    B, hq, hkv, max_seq, splits = 8, 32, 8, 4864, 6
    cos, sin = llama_cos_sin(TextSpec(head_dim=hd), max_seq)
    cos, sin = cos.to(bf).to(DEV), sin.to(bf).to(DEV)
    qkv = rnd(B, (hq + 2 * hkv) * hd, seed=95)
    kc0, vc0 = rnd(B, hkv, max_seq, hd, seed=96), rnd(B, hkv, max_seq, hd, seed=97)
    pos = torch.tensor([4672, 4673, 4700, 4799, 15, 16, 0, 3000], dtype=torch.int32, device=DEV)
    for b in range(B):
        kc0[b, :, int(pos[b]) + 1:] = float("nan")
        vc0[b, :, int(pos[b]) + 1:] = float("nan")
    kc1, vc1, q1 = kc0.clone(), vc0.clone(), qkv.clone()
    ops.llama_rope_kv_(q1, hq, hkv, hd, cos, sin, kc1, vc1, max_seq, pos_dev=pos, slot0=-1)
    kc2, vc2 = kc0.clone(), vc0.clone()
    got = ops.decode_attn_rope(qkv.clone(), kc2, vc2, pos, hq, hkv, hd, max_seq, cos, sin, splits=splits)
    torch.cuda.synchronize()
    assert torch.equal(kc1.nan_to_num(7.0), kc2.nan_to_num(7.0)) and torch.equal(vc1.nan_to_num(7.0), vc2.nan_to_num(7.0))
    assert bool(torch.isfinite(got.float()).all())
    for b in range(B):
        n = int(pos[b]) + 1
        q = q1[b, :hq * hd].view(hq, 1, hd).float()[None]
        k = kc1[b, :, :n].float().repeat_interleave(hq // hkv, 0)[None]
        v = vc1[b, :, :n].float().repeat_interleave(hq // hkv, 0)[None]
        ref = F.scaled_dot_product_attention(q, k, v)[0, :, 0].reshape(hq * hd)
        err = (got[b].float() - ref).abs().max().item()
        assert err <= 2e-2, (b, err)


# ---- R6+R7: assemble (embed + image scatter) and RoI replay, single and batched launch, bit-exact vs the oracle ----
def _assemble_case(spec, g, ncw, nch, n_masks, text=9):
    import numpy as np

    T = ncw * nch + 1
    ids = [128000] + [spec.image_token_id] * (T * spec.tokens_per_tile)
    boxes = {}
    for k in range(n_masks):
        ids += [int(v) for v in g.integers(0, 128000, size=text)] + [spec.crop_tokens_ids[k]] * spec.tokens_per_tile
        x0, y0 = g.uniform(0.0, 0.6, size=2)
        w, h = g.uniform(0.05, 0.39, size=2)
        boxes[str(spec.crop_tokens_ids[k])] = (float(x0), float(y0), float(x0 + w), float(y0 + h))
    ids += [int(v) for v in g.integers(0, 128000, size=text)] + [128009]
    return torch.tensor(ids, dtype=torch.long), boxes


@pytest.mark.parametrize("nreq,ncw,nch,n_masks", [(1, 2, 2, 1), (3, 2, 2, 2), (2, 3, 2, 3), (10, 1, 2, 1), (1, 4, 4, 5)])
def test_assemble_replay_batched_bit_exact(nreq, ncw, nch, n_masks):
    import numpy as np

    from grasp_any_region_b200.config import gar_tiny
    from grasp_any_region_b200.engine import roi_boxes_f32
    from oracle.gar_oracle import build_oracle

    spec = gar_tiny()
    oracle = build_oracle(spec, seed=0, dtype=torch.float32).to(bf)
    table = oracle.mllm.model.language_model.embed_tokens.weight.detach().to(DEV).contiguous()
    D = table.shape[1]
    tpt = int(math.isqrt(spec.tokens_per_tile))
    T = ncw * nch + 1
    g = np.random.default_rng(nreq * 100 + ncw * 10 + n_masks)
    cases = [_assemble_case(spec, g, ncw, nch, n_masks) for _ in range(nreq)]
    feats = rnd(nreq, T, spec.tokens_per_tile, D, seed=3)
    rois, has, grids = [], [], []
    for _, boxes in cases:
        rr, hh, scale = roi_boxes_f32(spec, boxes, ncw, nch)
        rois.append(rr)
        has.append(hh)
        grids.append((ncw, nch))
    ids_dev = [c[0].to(DEV) for c in cases]
    out, status = ops.assemble_embeds_batch(ids_dev, table, feats, list(spec.crop_tokens_ids), rois, has, scale, grids,
                                            tpt, spec.image_token_id)
    torch.cuda.synchronize()
    assert status[:, 0].tolist() == [0] * nreq and status[:, 1].tolist() == [T * spec.tokens_per_tile] * nreq
    for i, (ids, boxes) in enumerate(cases):
        want = oracle.assemble(ids[None], feats[i].cpu(), torch.tensor([[ncw, nch]]), [boxes])[0]
        assert torch.equal(out[i].cpu(), want), f"request {i}: assemble / RoI replay rows differ"
        # the single-request entry point is the same kernels with nreq = 1
        one, st1 = ops.assemble_embeds(ids_dev[i], table, feats[i], list(spec.crop_tokens_ids), rois[i], has[i], scale,
                                       ncw, nch, tpt, spec.image_token_id)
        assert torch.equal(one, out[i]) and int(st1[0]) == 0


@pytest.mark.parametrize("video,nreq,ncw,nch,n_masks", [(False, 1, 2, 2, 3), (False, 3, 3, 2, 2), (False, 2, 4, 4, 5),
                                                         (True, 1, 2, 2, 3), (True, 3, 2, 1, 3), (True, 2, 4, 2, 5)])
def test_assemble_plm_replay_modes_bit_exact(video, nreq, ncw, nch, n_masks):
    """the two replay branches of the vendored PLM's prepare_inputs_for_generation
    (modeling_perception_lm.py:648-852) as replay modes of the same kernels, bit-exact against the oracle restatement
    that tests/test_oracle_vs_reference.py pins to the reference's own run: GAR_REPLAY_MERGED_LAST keeps only the
    last crop token's replay, GAR_REPLAY_FRAME replays crop id k from tile (frame) k's own map."""
    import numpy as np

    from grasp_any_region_b200.config import gar_tiny
    from grasp_any_region_b200.engine import roi_boxes_f32
    from oracle.gar_oracle import build_oracle

    spec = gar_tiny()
    oracle = build_oracle(spec, seed=0, dtype=torch.float32).to(bf)
    table = oracle.mllm.model.language_model.embed_tokens.weight.detach().to(DEV).contiguous()
    D = table.shape[1]
    tpt = int(math.isqrt(spec.tokens_per_tile))
    T = ncw * nch + 1
    crop_ids = list(spec.crop_tokens_ids)  # frame f <-> crop id f in the per-frame mode (T >= n_masks here)
    g = np.random.default_rng(nreq * 100 + ncw * 10 + n_masks + 7 * video)
    cases = [_assemble_case(spec, g, ncw, nch, n_masks) for _ in range(nreq)]
    feats = rnd(nreq, T, spec.tokens_per_tile, D, seed=5)
    rois, has, grids = [], [], []
    for _, boxes in cases:
        rr, hh, scale = roi_boxes_f32(spec, boxes, ncw, nch, crop_ids, per_frame=video)
        rois.append(rr)
        has.append(hh)
        grids.append((ncw, nch))
    ids_dev = [c[0].to(DEV) for c in cases]
    mode = ops.REPLAY_FRAME if video else ops.REPLAY_MERGED_LAST
    out, status = ops.assemble_embeds_batch(ids_dev, table, feats, crop_ids, rois, has, scale, grids, tpt,
                                            spec.image_token_id, replay_mode=mode)
    plain, _ = ops.assemble_embeds_batch(ids_dev, table, feats, crop_ids, rois, has, scale, grids, tpt,
                                         spec.image_token_id)
    torch.cuda.synchronize()
    assert status[:, 0].tolist() == [0] * nreq
    for i, (ids, boxes) in enumerate(cases):
        want = oracle.assemble_plm(ids[None], feats[i].cpu(), torch.tensor([[ncw, nch]]), [boxes], crop_tokens=crop_ids,
                                   video=video, frame_crop_ids=crop_ids + [-1] * T)[0]
        assert torch.equal(out[i].cpu(), want), f"request {i}: rows differ from the reference's replay branch"
        if n_masks > 1:
            assert not torch.equal(out[i], plain[i]), "the mode must change the result when several regions are present"


@pytest.mark.parametrize("h,w,max_tiles", [(1024, 1024, 16), (600, 900, 6), (300, 520, 8), (448, 448, 4), (2000, 700, 16)])
def test_device_preprocessing_is_bit_identical_to_the_host_builder(h, w, max_tiles):
    """SURVEY §8f N1 device half: uint8 image + uint8 id map -> bf16 tiles on the GPU (ATen's uint8 antialiased bicubic
    passes, `_split`, rescale + normalise; NEAREST id-map tiling) equal the host builder's tensors -- which come from the
    same torchvision call the reference's image processor makes -- bit for bit."""
    import numpy as np

    from grasp_any_region_b200 import preprocess as pp
    from grasp_any_region_b200.config import gar_tiny
    from grasp_any_region_b200.preprocess_device import DevicePreprocessor

    spec = gar_tiny()
    rng = np.random.default_rng(h + w)
    img = torch.from_numpy(rng.integers(0, 256, size=(3, h, w), dtype=np.uint8))
    m0 = np.zeros((h, w), bool); m0[h // 10: h // 2, w // 8: w // 2] = True
    m1 = np.zeros((h, w), bool); m1[h // 3: h - 5, w // 3: w - 7] = True
    idmap = torch.from_numpy(pp.paint_id_map([m0, m1], spec.prompt_numbers))
    want_px, (ncw, nch) = pp.tile_image(img, spec.vision.img_size, max_tiles)
    want_mv = pp.tile_id_map(idmap, ncw, nch, spec.vision.img_size)
    pre = DevicePreprocessor(spec, DEV)
    px, mv, grid = pre(img.to(DEV), idmap.to(DEV), max_tiles)
    torch.cuda.synchronize()
    assert grid == (ncw, nch) and px.dtype == bf and px.shape == want_px.shape and mv.shape == want_mv.shape
    assert torch.equal(px.cpu(), want_px.to(bf)), "pixel tiles differ from the host (torchvision) path"
    assert torch.equal(mv.cpu(), want_mv.to(bf)), "id-map tiles differ from the host path"
