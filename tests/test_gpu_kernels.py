"""GPU parity tests of every C-ABI kernel against the oracle's expressions (oracle/wan_oracle.py) -- bit-level for
the elementwise kernels up to fp32 reduction order, bf16-rounding-level for GEMM / attention -- plus
size-independent properties at the full BASELINE sizes and the edge cases (ragged, tails, empty)."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import wan_oracle as O
from video_styler_b200 import _lib, ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _no_timeouts():
    f = _lib.debug_flags()
    assert f["timeouts"] == 0, f


def _close(got, ref, rel, mism=None):
    m = O.parity_metrics(got.float(), ref.float())
    assert m["rel_l2"] <= rel, m
    assert not torch.isnan(got.float()).any()
    if mism is not None:
        frac = float((got.cpu() != ref.cpu()).double().mean())
        assert frac <= mism, (frac, m)


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 1e-4), (torch.float32, 1e-6)])
@pytest.mark.parametrize("n,d", [(72, 256), (45, 512), (1280, 1536), (777, 5120)])
def test_ln_modulate_matches_oracle(n, d, dtype, tol):
    g = torch.Generator().manual_seed(n + d)
    x = (torch.randn(n, d, generator=g) * 2 + 0.3).to(dtype)
    sh, sc = (torch.randn(d, generator=g) * 0.5).to(dtype), (torch.randn(d, generator=g) * 0.5).to(dtype)
    w, b = (1 + 0.1 * torch.randn(d, generator=g)).to(dtype), (0.1 * torch.randn(d, generator=g)).to(dtype)
    ref_mod = O.layer_norm(x, 1e-6) * (1 + sc) + sh          # wan_video_dit.py:225 in the tensor dtype
    ref_aff = O.layer_norm(x, 1e-6, w, b)
    xd = x.to(DEV)
    _close(ops.ln_modulate(xd, sh.to(DEV), sc.to(DEV), eps=1e-6), ref_mod, tol, 1e-3 if dtype == torch.bfloat16 else None)
    _close(ops.ln_modulate(xd, weight=w.to(DEV), bias=b.to(DEV), eps=1e-6), ref_aff, tol, 1e-3 if dtype == torch.bfloat16 else None)
    _close(ops.ln_modulate(xd, eps=1e-6), O.layer_norm(x, 1e-6), tol)
    # strided input view (column slice of a wider buffer)
    wide = torch.zeros(n, 2 * d, dtype=dtype, device=DEV)
    wide[:, d:] = xd
    _close(ops.ln_modulate(wide[:, d:], sh.to(DEV), sc.to(DEV), eps=1e-6), ref_mod, tol)


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 2e-4), (torch.float32, 1e-6)])
@pytest.mark.parametrize("grid,d,offset", [((3, 4, 6), 256, 0), ((5, 16, 16), 1536, 0), ((2, 30, 52), 5120, 0), ((4, 6, 10), 512, 120)])
def test_qk_rmsnorm_rope_matches_oracle(grid, d, offset, dtype, tol):
    gf, gh, gw = grid
    n_total = gf * gh * gw
    n = n_total - offset
    heads = d // 128
    g = torch.Generator().manual_seed(d + n)
    q, k = torch.randn(n, d, generator=g).to(dtype), torch.randn(n, d, generator=g).to(dtype)
    wq, wk = (1 + 0.1 * torch.randn(d, generator=g)).to(dtype), (1 + 0.1 * torch.randn(d, generator=g)).to(dtype)
    freqs = O.rope_freqs(128, gf, gh, gw)[offset:]
    ref_q = O.rope_apply(O.rms_norm(q[None], wq, 1e-6), freqs, heads)[0]
    ref_k = O.rope_apply(O.rms_norm(k[None], wk, 1e-6), freqs, heads)[0]
    table = ops.make_rope_table(O.rope_tables_3d(128), DEV)
    buf = torch.zeros(n, 3 * d, dtype=dtype, device=DEV)          # the fused q|k|v layout of the engine
    buf[:, :d], buf[:, d:2 * d] = q.to(DEV), k.to(DEV)
    ops.qk_rmsnorm_rope(buf[:, :d], buf[:, d:2 * d], wq.to(DEV), wk.to(DEV), 1e-6, table, grid, offset)
    mism = 1e-3 if dtype == torch.bfloat16 else None
    _close(buf[:, :d], ref_q, tol, mism)
    _close(buf[:, d:2 * d], ref_k, tol, mism)
    assert float(buf[:, 2 * d:].abs().max()) == 0.0
    # cross-attention form: no rope, q only
    qo, _ = ops.qk_rmsnorm_rope(q.to(DEV).clone(), None, wq.to(DEV), None, 1e-6)
    _close(qo, O.rms_norm(q, wq, 1e-6), tol, mism)


def test_rope_frame_ids_variant():
    gf, gh, gw, d = 3, 4, 6, 256
    n = gf * gh * gw
    ids = torch.tensor([0, 9, 33])
    q = torch.randn(n, d).bfloat16()
    w = torch.ones(d).bfloat16()
    ref = O.rope_apply(O.rms_norm(q[None], w, 1e-6), O.rope_freqs(128, gf, gh, gw, rope_indices=ids), 2)[0]
    table = ops.make_rope_table(O.rope_tables_3d(128), DEV)
    got, _ = ops.qk_rmsnorm_rope(q.to(DEV), None, w.to(DEV), None, 1e-6, table, (gf, gh, gw), 0, ids.int().to(DEV))
    _close(got, ref, 2e-4, 1e-3)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_rope_per_token_table_and_padded_rows(dtype):
    """(i) per-token mode: the reference's complex (N, 1, 64) freqs tensor as a (N, 64, 2) table gives the same bits
    as the table + grid mode; (ii) rows past the token grid -- the zero padding of the last Ulysses shard
    (wan_video_new.py:1414-1416; a ragged N on the LAST rank) -- are accepted, leave the real rows untouched and
    produce finite values (they are never attended)."""
    gf, gh, gw, d = 3, 5, 7, 512
    n = gf * gh * gw
    g = torch.Generator().manual_seed(3)
    q, k = torch.randn(n, d, generator=g).to(dtype), torch.randn(n, d, generator=g).to(dtype)
    w = (1 + 0.1 * torch.randn(d, generator=g)).to(dtype)
    table = ops.make_rope_table(O.rope_tables_3d(128), DEV)
    a_q, a_k = ops.qk_rmsnorm_rope(q.to(DEV), k.to(DEV), w.to(DEV), w.to(DEV), 1e-6, table, (gf, gh, gw), 0)
    per_tok = ops.rope_table_from_freqs(O.rope_freqs(128, gf, gh, gw), DEV)
    b_q, b_k = ops.qk_rmsnorm_rope(q.to(DEV), k.to(DEV), w.to(DEV), w.to(DEV), 1e-6, per_tok, (0, 0, 0), 0)
    assert torch.equal(a_q, b_q) and torch.equal(a_k, b_k)
    # last shard of a 2-rank split of n = 105 tokens: n_loc = 53, rank 1 holds rows [53, 105) + 1 pad row
    n_loc = 53
    shard = torch.zeros(n_loc, d, dtype=dtype)
    shard[:n - n_loc] = q[n_loc:]
    shard[n - n_loc:] = 7.0                                    # stale data in the pad row
    c_q, _ = ops.qk_rmsnorm_rope(shard.to(DEV), None, w.to(DEV), None, 1e-6, table, (gf, gh, gw), n_loc)
    assert torch.equal(c_q[:n - n_loc], a_q[n_loc:]) and torch.isfinite(c_q.float()).all()
    with pytest.raises(_lib.WvdError):
        ops.qk_rmsnorm_rope(shard.to(DEV), None, w.to(DEV), None, 1e-6, table, (gf, gh, gw), n + 1)     # offset outside the grid
    _no_timeouts()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_residual_forms(dtype):
    x, y, gt = torch.randn(300, 512).to(dtype), torch.randn(300, 512).to(dtype), torch.randn(512).to(dtype)
    assert torch.equal(ops.scale_add(x.to(DEV), y.to(DEV), 0.75).cpu(), x + y * 0.75)
    assert torch.equal(ops.gate_residual(x.to(DEV), gt.to(DEV), y.to(DEV)).cpu(), x + gt * y)


def _gemm_ref(x, w, b, epi, gate, res):
    y = F.linear(x.float(), w.float(), None if b is None else b.float()).to(x.dtype)     # F.linear rounding
    if epi == ops.EPI_BIAS_GELU:
        y = F.gelu(y, approximate="tanh")
    elif epi == ops.EPI_BIAS_RES:
        y = res + y
    elif epi == ops.EPI_BIAS_GATE_RES:
        y = res + gate * y
    return y


@pytest.fixture(params=[_lib.GEMM_1CTA, _lib.GEMM_2CTA, _lib.GEMM_2CTA_M512], ids=["one_cta_tiles", "cta_pair_tiles", "cta_pair_m512_tiles"])
def gemm_variant(request):
    """Run the test on ALL GEMM variants (128 x 256 tiles per CTA; 256 x 256 and 512 x 256 tiles per CTA pair, cta_group::2)."""
    return request.param


@pytest.mark.parametrize("m,n,k", [(1, 8, 8), (72, 256, 256), (129, 264, 72), (200, 512, 320), (512, 1536, 1536),
                                    (1280, 8960, 1536), (300, 1536, 8960), (520, 64, 512), (385, 520, 64)])
def test_gemm_bf16_all_epilogues(m, n, k, gemm_variant):
    g = torch.Generator().manual_seed(m * 7 + n)
    x = torch.randn(m, k, generator=g).bfloat16()
    w = (torch.randn(n, k, generator=g) / math.sqrt(k)).bfloat16()
    b = (torch.randn(n, generator=g) * 0.1).bfloat16()
    gate, res = torch.randn(n, generator=g).bfloat16(), torch.randn(m, n, generator=g).bfloat16()
    xd, wd, bd, gd, rd = (t.to(DEV) for t in (x, w, b, gate, res))
    for epi in (ops.EPI_BIAS, ops.EPI_BIAS_GELU, ops.EPI_BIAS_RES, ops.EPI_BIAS_GATE_RES):
        out = ops.linear(xd, wd, bd, epi, gd if epi == ops.EPI_BIAS_GATE_RES else None,
                         rd if epi in (ops.EPI_BIAS_RES, ops.EPI_BIAS_GATE_RES) else None, variant=gemm_variant)
        _close(out, _gemm_ref(x, w, b, epi, gate, res), 5e-4, 1e-2)       # 1-ulp flips from fp32 summation order
    _close(ops.linear(xd, wd, variant=gemm_variant), _gemm_ref(x, w, None, 0, None, None), 5e-4, 1e-2)          # no bias
    r2 = rd.clone()                                                                      # in-place residual stream
    ops.linear(xd, wd, bd, ops.EPI_BIAS_GATE_RES, gd, r2, out=r2, variant=gemm_variant)
    _close(r2, _gemm_ref(x, w, b, ops.EPI_BIAS_GATE_RES, gate, res), 5e-4, 1e-2)
    # output into a column slice of a wider buffer (the q|k|v layout): neighbours untouched (TMA store clipping)
    wide = torch.full((m, n + 16), 7.0, dtype=torch.bfloat16, device=DEV)
    ops.linear(xd, wd, bd, out=wide[:, 8:8 + n], variant=gemm_variant)
    _close(wide[:, 8:8 + n], _gemm_ref(x, w, b, 0, None, None), 5e-4, 1e-2)
    assert float((wide[:, :8] - 7).abs().max()) == 0 and float((wide[:, 8 + n:] - 7).abs().max()) == 0
    _no_timeouts()


@pytest.mark.parametrize("m,n,k,g", [(300, 256, 512, 3), (1000, 1536, 1536, 3), (129, 320, 72, 2), (3705, 5120, 512, 3)])
def test_gemm_grouped_equals_separate_launches(m, n, k, g, gemm_variant):
    """The grouped q|k|v launch (one A, up to three weights, column slices of one output) is bit-identical to the
    separate launches and matches F.linear."""
    gen = torch.Generator().manual_seed(m + n + g)
    x = torch.randn(m, k, generator=gen).bfloat16().to(DEV)
    ws = [(torch.randn(n, k, generator=gen) / math.sqrt(k)).bfloat16().to(DEV) for _ in range(g)]
    bs = [(torch.randn(n, generator=gen) * 0.1).bfloat16().to(DEV) for _ in range(g)]
    out = torch.full((m, g * n + 8), 3.0, dtype=torch.bfloat16, device=DEV)
    ops.linear_grouped(x, ws, bs, out=out[:, :g * n], variant=gemm_variant)
    for i in range(g):
        sep = ops.linear(x, ws[i], bs[i], variant=gemm_variant)
        assert torch.equal(out[:, i * n:(i + 1) * n], sep), i
        _close(sep, _gemm_ref(x.cpu(), ws[i].cpu(), bs[i].cpu(), 0, None, None), 5e-4, 1e-2)
    assert float((out[:, g * n:] - 3).abs().max()) == 0
    # a missing bias in one group
    ops.linear_grouped(x, ws, [bs[0]] + [None] * (g - 1), out=out[:, :g * n], variant=gemm_variant)
    assert torch.equal(out[:, (g - 1) * n:g * n], ops.linear(x, ws[g - 1], None, variant=gemm_variant))
    _no_timeouts()
    _no_timeouts()


def test_gemm_exact_on_integers_and_strided_output():
    """Small integers are exact in bf16 and in the fp32 accumulator: any tile/descriptor/swizzle error is visible."""
    g = torch.Generator().manual_seed(0)
    for (m, n, k) in [(128, 256, 64), (256, 512, 256), (384, 768, 192), (130, 520, 136)]:
        x = torch.randint(-2, 3, (m, k), generator=g).bfloat16()
        w = torch.randint(-2, 3, (n, k), generator=g).bfloat16()
        buf = torch.zeros(m, 3 * n, dtype=torch.bfloat16, device=DEV)
        ops.linear(x.to(DEV), w.to(DEV), out=buf[:, n:2 * n])
        assert torch.equal(buf[:, n:2 * n].cpu(), (x.float() @ w.float().t()).bfloat16())
        assert float(buf[:, :n].abs().max()) == 0 and float(buf[:, 2 * n:].abs().max()) == 0
    _no_timeouts()


def test_gemm_linearity_at_full_c3_shape(gemm_variant):
    """Size-independent property at the BASELINE shape (29,640 x 5120 x 5120): W(2x) = 2 W(x) exactly in bf16
    (scaling by 2 is exact), and W(x) on a row subset equals the subset of W(x)."""
    m, n, k = 29640, 5120, 5120
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(m, k, device=DEV, generator=g).bfloat16()
    w = (torch.randn(n, k, device=DEV, generator=g) / math.sqrt(k)).bfloat16()
    y = ops.linear(x, w, variant=gemm_variant)
    y2 = ops.linear(x * 2, w, variant=gemm_variant)
    assert torch.equal(y2, y * 2)
    assert torch.equal(y, ops.linear(x, w, variant=gemm_variant % 3 + 1))      # all variants accumulate in the same order
    rows = torch.tensor([0, 127, 128, 14820, 29567, 29568, 29639], device=DEV)
    ysub = ops.linear(x[rows].contiguous(), w, variant=gemm_variant)
    assert torch.equal(ysub, y[rows])
    ref = F.linear(x[rows].float(), w.float()).bfloat16()
    _close(ysub, ref, 5e-4, 5e-3)
    _no_timeouts()


def _attn_ref(q, k, v, h, dtype=torch.float32):
    return O.attention(q[None].to(dtype), k[None].to(dtype), v[None].to(dtype), h)[0]


@pytest.fixture(params=[1, 2, 3, 4, 5], ids=["two_tile_kernel", "cta_pair_kernel", "cta_group2_kernel", "one_tile_kernel",
                                          "cta_group2_persistent_kernel"])
def attn_kernel(request):
    """Run the test on ALL bf16 attention kernels (the default dispatch picks by key length)."""
    return request.param


@pytest.mark.parametrize("sq,sk,h", [(1, 1, 1), (128, 128, 1), (72, 72, 2), (256, 256, 2), (300, 512, 2), (257, 129, 3),
                                      (1280, 1280, 12), (1000, 777, 3), (640, 2100, 5)])
def test_attention_matches_oracle(sq, sk, h, attn_kernel):
    g = torch.Generator().manual_seed(sq + sk)
    q, k, v = (torch.randn(s, h * 128, generator=g).bfloat16() for s in (sq, sk, sk))
    out = ops.attention(q.to(DEV), k.to(DEV), v.to(DEV), h, kernel=attn_kernel)
    _no_timeouts()
    _close(out, _attn_ref(q, k, v, h), 4e-3)            # bf16 P / bf16 output rounding (same points as FA2)


def test_attention_repeated_launches_are_bit_identical(attn_kernel):
    """Warp-specialised pipelines fail as races: the same launch 40 times, with other work in between, must give the
    same bits every time and never trip the in-kernel watchdog (a parity-aliasing deadlock of the CTA-pair kernel
    showed up in ~1 of 100 launches before its P hand-over barriers were indexed by S buffer)."""
    n, h = 8192, 8
    g = torch.Generator(device=DEV).manual_seed(11)
    qkv = torch.randn(n, 3 * h * 128, device=DEV, generator=g).bfloat16()
    d = h * 128
    junk = torch.empty(160 << 20, dtype=torch.uint8, device=DEV)
    ref = None
    for _ in range(40):
        junk.add_(1)
        out = ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], h, kernel=attn_kernel)
        ref = out.clone() if ref is None else ref
        assert torch.equal(out, ref)
    _no_timeouts()


def test_attention_reads_fused_qkv_views_and_large_scores(attn_kernel):
    """q|k|v column slices of one buffer (the engine's layout); large-magnitude scores exercise the lazy rescale."""
    sq, h = 640, 2
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(sq, 3 * h * 128, generator=g)
    qkv[:, :h * 128] *= 4.0                             # sharp softmax, running max keeps growing
    qkv[300:, h * 128:2 * h * 128] *= 3.0
    qkv = qkv.bfloat16()
    d = h * 128
    buf = qkv.to(DEV)
    out = ops.attention(buf[:, :d], buf[:, d:2 * d], buf[:, 2 * d:], h, kernel=attn_kernel)
    _no_timeouts()
    _close(out, _attn_ref(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], h), 6e-3)


@pytest.mark.parametrize("n", [29640, 75600])
def test_attention_properties_at_full_length(attn_kernel, n):
    """29,640 tokens (config c3) and 75,600 (config c5), 2 heads: (i) V = const -> output = const exactly up to bf16;
    (ii) key/value permutation invariance; (iii) a row subset equals SDPA in fp32."""
    h = 2
    g = torch.Generator(device=DEV).manual_seed(2)
    q = torch.randn(n, h * 128, device=DEV, generator=g).bfloat16()
    k = torch.randn(n, h * 128, device=DEV, generator=g).bfloat16()
    v = torch.randn(n, h * 128, device=DEV, generator=g).bfloat16()
    ones = torch.full_like(v, 0.5)
    o1 = ops.attention(q, k, ones, h, kernel=attn_kernel)
    assert float((o1.float() - 0.5).abs().max()) <= 0.5 * 2 ** -7
    out = ops.attention(q, k, v, h, kernel=attn_kernel)
    perm = torch.randperm(n, device=DEV, generator=g)
    outp = ops.attention(q, k[perm].contiguous(), v[perm].contiguous(), h, kernel=attn_kernel)
    assert O.parity_metrics(outp, out)["rel_l2"] <= 4e-3
    rows = torch.arange(0, n, 997, device=DEV)
    ref = _attn_ref(q[rows].cpu(), k.cpu(), v.cpu(), h) if False else \
        O.attention(q[rows][None].float(), k[None].float(), v[None].float(), h)[0]
    _close(out[rows], ref.cpu(), 4e-3)
    _no_timeouts()


def test_fp32_kernels_match_oracle():
    g = torch.Generator().manual_seed(9)
    x, w, b = torch.randn(200, 320, generator=g), torch.randn(264, 320, generator=g) / 18, torch.randn(264, generator=g)
    res, gate = torch.randn(200, 264, generator=g), torch.randn(264, generator=g)
    for epi, ref in [(ops.EPI_BIAS, F.linear(x, w, b)), (ops.EPI_BIAS_GELU, F.gelu(F.linear(x, w, b), approximate="tanh")),
                     (ops.EPI_BIAS_RES, res + F.linear(x, w, b)), (ops.EPI_BIAS_GATE_RES, res + gate * F.linear(x, w, b))]:
        out = ops.linear(x.to(DEV), w.to(DEV), b.to(DEV), epi, gate.to(DEV) if epi == 3 else None,
                         res.to(DEV) if epi >= 2 else None)
        _close(out, ref, 2e-6)
    q, k, v = (torch.randn(s, 256, generator=g) for s in (300, 200, 200))
    _close(ops.attention(q.to(DEV), k.to(DEV), v.to(DEV), 2), _attn_ref(q, k, v, 2), 5e-6)


def test_ulysses_layout_kernels():
    n, heads, world = 37, 4, 2
    qkv = torch.randn(n, 3 * heads * 128).bfloat16()
    from tests import cpu_backend
    got = ops.ulysses_pack_qkv(qkv.to(DEV), heads, world)
    assert torch.equal(got.cpu(), cpu_backend.ulysses_pack_qkv(qkv, heads, world))
    recv = torch.randn(world, n, (heads // world) * 128).bfloat16()
    assert torch.equal(ops.ulysses_unpack_out(recv.to(DEV), heads, world).cpu(), cpu_backend.ulysses_unpack_out(recv, heads, world))


@pytest.mark.parametrize("world,heads,n_loc,pad", [(2, 4, 37, 4), (2, 4, 1100, 3), (4, 8, 600, 0)])
def test_ulysses_peer_scatter_kernels_match_the_collective_layouts(world, heads, n_loc, pad):
    """The fused exchange (peer stores) against the pack / unpack layouts, with the 'peers' being local buffers: what
    rank r scatters must land where pack + all_to_all would have put it, and attention_scatter must leave every query
    row where attention + all_to_all + unpack would have (ragged token count included)."""
    from tests import cpu_backend
    hl, w = heads // world, (heads // world) * 128
    g = torch.Generator().manual_seed(5)
    qkv = [torch.randn(n_loc, 3 * heads * 128, generator=g).bfloat16() for _ in range(world)]
    recv = [torch.zeros(world * n_loc, 3 * w, dtype=torch.bfloat16, device=DEV) for _ in range(world)]
    for r in range(world):
        ops.ulysses_scatter_qkv(qkv[r].to(DEV), heads, [t.data_ptr() for t in recv], r)
    torch.cuda.synchronize()
    packed = [cpu_backend.ulysses_pack_qkv(q, heads, world) for q in qkv]          # (world, n_loc, 3, hl, 128) per source
    for d in range(world):
        want = torch.stack([packed[src][d] for src in range(world)]).reshape(world * n_loc, 3 * w)   # all_to_all_single
        assert torch.equal(recv[d].cpu(), want)
    # the same receive buffers filled by the RoPE kernel's fused stores (q, k) + the v-only scatter: bit-identical to
    # RMSNorm+RoPE in place followed by the full scatter
    table = ops.make_rope_table(O.rope_tables_3d(128), DEV)
    grid = (world, n_loc // 4 + 1, 4)                                   # any grid with >= world * n_loc - pad tokens
    wq = (1 + 0.1 * torch.randn(heads * 128, generator=g)).bfloat16().to(DEV)
    wk = (1 + 0.1 * torch.randn(heads * 128, generator=g)).bfloat16().to(DEV)
    recv_a = [torch.zeros_like(t) for t in recv]
    recv_b = [torch.zeros_like(t) for t in recv]
    d_ = heads * 128
    for r in range(world):
        buf = qkv[r].to(DEV)
        ops.qk_rmsnorm_rope_scatter(buf[:, :d_], buf[:, d_:2 * d_], wq, wk, 1e-6, table, grid, r * n_loc, None,
                                    [t.data_ptr() for t in recv_b], r)
        ops.ulysses_scatter_v(buf, heads, [t.data_ptr() for t in recv_b], r)
        assert torch.equal(buf.cpu(), qkv[r])                           # q, k left untouched
        ops.qk_rmsnorm_rope(buf[:, :d_], buf[:, d_:2 * d_], wq, wk, 1e-6, table, grid, r * n_loc)
        ops.ulysses_scatter_qkv(buf, heads, [t.data_ptr() for t in recv_a], r)
    torch.cuda.synchronize()
    for d in range(world):
        assert torch.equal(recv_a[d], recv_b[d]), d
    # return trip: the last shard is padded by `pad` rows; every rank attends its heads over all real tokens (the
    # 2,197- and 2,400-token cases take the long-sequence CTA-pair kernel's scatter epilogue)
    n = world * n_loc - pad
    outs = [torch.zeros(n_loc, heads * 128, dtype=torch.bfloat16, device=DEV) for _ in range(world)]
    for r in range(world):
        rv = recv[r]
        ops.attention_scatter(rv[:n, :w], rv[:n, w:2 * w], rv[:n, 2 * w:], hl, [t.data_ptr() for t in outs], heads * 128,
                              n_loc, r * w)
    torch.cuda.synchronize()
    for r in range(world):
        rv = recv[r]
        full = torch.zeros(world * n_loc, w, dtype=torch.bfloat16, device=DEV)
        ops.attention(rv[:n, :w], rv[:n, w:2 * w], rv[:n, 2 * w:], hl, out=full[:n])
        # ... and the plain kernel itself against the fp32 oracle on the same operands (not only against itself)
        _close(full[:n], _attn_ref(rv[:n, :w].cpu(), rv[:n, w:2 * w].cpu(), rv[:n, 2 * w:].cpu(), hl), 4e-3)
        for d in range(world):
            assert torch.equal(outs[d][:, r * w:(r + 1) * w], full[d * n_loc:(d + 1) * n_loc]), (r, d)
    _no_timeouts()


def test_empty_and_invalid_inputs():
    x = torch.empty(0, 256, dtype=torch.bfloat16, device=DEV)
    assert ops.ln_modulate(x, eps=1e-6).shape == (0, 256)
    assert ops.linear(x, torch.randn(64, 256, device=DEV).bfloat16()).shape == (0, 64)
    with pytest.raises(_lib.WvdError):
        ops.linear(torch.randn(4, 12, device=DEV).bfloat16(), torch.randn(8, 12, device=DEV).bfloat16())    # K % 8
    with pytest.raises(_lib.WvdError):
        ops.attention(torch.randn(4, 64, device=DEV).bfloat16(), torch.randn(4, 64, device=DEV).bfloat16(),
                      torch.randn(4, 64, device=DEV).bfloat16(), 1)                                          # head_dim 64
    with pytest.raises(_lib.WvdError):
        ops.ln_modulate(torch.randn(4, 256, device=DEV).half(), eps=1e-6)                                    # fp16
