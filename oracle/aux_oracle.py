"""CPU restatement of the two "next" rows either side of the DiT  --  TEST INFRASTRUCTURE.

  umT5 text encoder      diffsynth/models/wan_video_text_encoder.py   (WanTextEncoder.forward and its blocks)
  keyframe editor step   diffsynth/pipelines/wan_video_editor.py      (compute_velocity_correction, the loop body)

Plain torch over a flat state dict, each function citing the reference lines it follows.  Pinned:
oracle/make_golden_aux.py runs the REAL reference classes on the same seeds and commits their outputs under
tests/golden/ (t5_tiny.pt, editor_step.pt); tests/test_oracle.py holds these functions to those.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import math

import torch
import torch.nn.functional as F

T5_CONFIGS = {
    "tiny": dict(vocab=1000, dim=256, dim_attn=256, dim_ffn=640, num_heads=4, num_layers=2, num_buckets=32, shared_pos=False),
    "small": dict(vocab=4096, dim=512, dim_attn=512, dim_ffn=1280, num_heads=8, num_layers=3, num_buckets=32, shared_pos=False),
    "umt5-xxl": dict(vocab=256384, dim=4096, dim_attn=4096, dim_ffn=10240, num_heads=64, num_layers=24, num_buckets=32,
                     shared_pos=False),
}


def t5_param_shapes(cfg):
    d, da, f, h, nb = cfg["dim"], cfg["dim_attn"], cfg["dim_ffn"], cfg["num_heads"], cfg["num_buckets"]
    shapes = {"token_embedding.weight": (cfg["vocab"], d), "norm.weight": (d,)}
    if cfg["shared_pos"]:
        shapes["pos_embedding.embedding.weight"] = (nb, h)
    for i in range(cfg["num_layers"]):
        p = f"blocks.{i}."
        shapes.update({p + "norm1.weight": (d,), p + "norm2.weight": (d,), p + "attn.q.weight": (da, d),
                       p + "attn.k.weight": (da, d), p + "attn.v.weight": (da, d), p + "attn.o.weight": (d, da),
                       p + "ffn.gate.0.weight": (f, d), p + "ffn.fc1.weight": (f, d), p + "ffn.fc2.weight": (d, f)})
        if not cfg["shared_pos"]:
            shapes[p + "pos_embedding.embedding.weight"] = (nb, h)
    return shapes


def make_t5_state_dict(cfg, seed=0, dtype=torch.float32, device="cpu"):
    """Synthetic weights in the spirit of the reference's init_weights (wan_video_text_encoder.py:177-194: normal, a
    standard deviation per layer kind) but scaled so that every part of the block matters numerically: k, v, gate, fc1
    ~ N(0, 1/fan_in); q a quarter of that (T5 attention has no 1/sqrt(d) factor: logits then have a standard deviation
    of ~2); o and fc2 half of 1/sqrt(fan_in) (a residual stream that grows slowly over 24 layers); norm weights
    1 + 0.1 N(0,1); token embedding N(0,1); bucket embedding N(0, 0.5) (a position bias that matters in the softmax).
    ``device``: draw on that device (the 4.6-B-parameter umT5-XXL materialises in seconds on a GPU); the golden fixtures
    use the CPU stream."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    for name, shape in t5_param_shapes(cfg).items():
        r = torch.randn(shape, generator=g, device=device)
        if name.endswith("norm1.weight") or name.endswith("norm2.weight") or name == "norm.weight":
            t = 1.0 + 0.1 * r
        elif name == "token_embedding.weight":
            t = r
        elif name.endswith("embedding.weight"):
            t = 0.5 * r
        else:
            gain = 0.25 if name.endswith("attn.q.weight") else (0.5 if name.endswith(("attn.o.weight", "ffn.fc2.weight")) else 1.0)
            t = r * (gain / math.sqrt(shape[1]))
        sd[name] = t.to(dtype)
    return sd


def make_t5_inputs(cfg, length=40, valid=23, seed=1):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1, cfg["vocab"], (1, length), generator=g)
    ids[:, valid:] = 0
    mask = torch.zeros(1, length, dtype=torch.long)
    mask[:, :valid] = 1
    return ids, mask


def t5_layer_norm(x, weight, eps=1e-6):
    """T5LayerNorm.forward (wan_video_text_encoder.py:30-35)."""
    x = x * torch.rsqrt(x.float().pow(2).mean(dim=-1, keepdim=True) + eps)
    if weight.dtype in (torch.float16, torch.bfloat16):
        x = x.type_as(weight)
    return weight * x


def t5_bucket(rel_pos, num_buckets=32, max_dist=128):
    """T5RelativeEmbedding._relative_position_bucket, bidirectional (:155-175)."""
    nb = num_buckets // 2
    out = (rel_pos > 0).long() * nb
    rp = torch.abs(rel_pos)
    max_exact = nb // 2
    large = max_exact + (torch.log(rp.float() / max_exact) / math.log(max_dist / max_exact) * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return out + torch.where(rp < max_exact, rp, large)


def t5_pos_bias(emb_weight, lq, lk, num_buckets=32):
    """T5RelativeEmbedding.forward (:141-153) -> (1, N, Lq, Lk)."""
    dev = emb_weight.device
    rel = torch.arange(lk, device=dev).unsqueeze(0) - torch.arange(lq, device=dev).unsqueeze(1)
    return F.embedding(t5_bucket(rel, num_buckets), emb_weight).permute(2, 0, 1).unsqueeze(0).contiguous()


def t5_gelu(x):
    """GELU.forward (:16-20), evaluated op by op in x's dtype like the reference."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def t5_attention(x, sd, p, num_heads, mask, pos_bias):
    """T5Attention.forward (:55-90): no scaling, additive bias, masked_fill(finfo.min), fp32 softmax."""
    b, n = x.size(0), num_heads
    q = F.linear(x, sd[p + "q.weight"]).view(b, -1, n, sd[p + "q.weight"].shape[0] // n)
    k = F.linear(x, sd[p + "k.weight"]).view(b, -1, n, q.shape[-1])
    v = F.linear(x, sd[p + "v.weight"]).view(b, -1, n, q.shape[-1])
    attn_bias = x.new_zeros(b, n, q.size(1), k.size(1))
    if pos_bias is not None:
        attn_bias += pos_bias
    if mask is not None:
        m = mask.view(b, 1, 1, -1)
        attn_bias.masked_fill_(m == 0, torch.finfo(x.dtype).min)
    attn = torch.einsum("binc,bjnc->bnij", q, k) + attn_bias
    attn = F.softmax(attn.float(), dim=-1).type_as(attn)
    y = torch.einsum("bnij,bjnc->binc", attn, v).reshape(b, -1, n * q.shape[-1])
    return F.linear(y, sd[p + "o.weight"])


def t5_encoder(sd, cfg, ids, mask=None):
    """WanTextEncoder.forward (:233-243) in eval mode (dropout = identity), per-block position bias (shared_pos False)."""
    x = F.embedding(ids, sd["token_embedding.weight"])
    l = x.size(1)
    e = t5_pos_bias(sd["pos_embedding.embedding.weight"], l, l, cfg["num_buckets"]).to(x.dtype) if cfg["shared_pos"] else None
    for i in range(cfg["num_layers"]):
        p = f"blocks.{i}."
        pb = e if cfg["shared_pos"] else t5_pos_bias(sd[p + "pos_embedding.embedding.weight"], l, l, cfg["num_buckets"]).to(x.dtype)
        x = x + t5_attention(t5_layer_norm(x, sd[p + "norm1.weight"]), sd, p + "attn.", cfg["num_heads"], mask, pb)   # :136
        h = t5_layer_norm(x, sd[p + "norm2.weight"])
        x = x + F.linear(F.linear(h, sd[p + "ffn.fc1.weight"]) * t5_gelu(F.linear(h, sd[p + "ffn.gate.0.weight"])),
                         sd[p + "ffn.fc2.weight"])                                                                    # :108-112, 137
    return t5_layer_norm(x, sd["norm.weight"])


def encode_prompt(sd, cfg, ids, mask):
    """WanPrompter.encode_prompt after the tokenizer (prompters/wan_prompter.py:105-109)."""
    seq_lens = mask.gt(0).sum(dim=1).long()
    emb = t5_encoder(sd, cfg, ids, mask)
    for _, v in enumerate(seq_lens):
        emb[:, v:] = 0
    return emb


# ---------------------------------------------------------------------------------------------------------------
# keyframe editor (wan_video_editor.py)
# ---------------------------------------------------------------------------------------------------------------
def velocity_correction(z_main, z_edit, v_main, v_edit, keyframe_indices, dt, alpha=10.0, beta=0.0):
    """compute_velocity_correction (:107-165)."""
    idx = list(keyframe_indices)
    z_diff = z_main[:, :, idx] - z_edit
    v_diff = v_main[:, :, idx] - v_edit
    r_k = z_diff - v_diff * dt
    correction = alpha * r_k
    v_main_c = v_main.clone()
    v_main_c[:, :, idx] += correction
    v_edit_c = v_edit - beta * correction if beta > 0 else v_edit
    return v_main_c, v_edit_c


def editor_step(z_main, z_edit, v_posi, v_nega, keyframe_indices, cfg_scale, dt, alpha, beta, dsigma):
    """The loop body after the DiT calls (:366-390): CFG, split, correction, Euler step of both latent sets
    (flow_match.py:72-82: sample + model_output * (sigma_next - sigma))."""
    v = v_posi if v_nega is None else v_nega + cfg_scale * (v_posi - v_nega)
    v_main, v_edit = torch.split(v, [z_main.shape[2], z_edit.shape[2]], dim=2)
    v_main_c, v_edit_c = velocity_correction(z_main, z_edit, v_main, v_edit, keyframe_indices, dt, alpha, beta)
    ds = torch.tensor(dsigma, dtype=torch.float32)
    return z_main + v_main_c * ds, z_edit + v_edit_c * ds


def make_editor_inputs(shape=(1, 16, 7, 8, 12), keyframes=(0, 3, 6), seed=5, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    b, c, t, h, w = shape
    k = len(keyframes)
    r = lambda *s: torch.randn(*s, generator=g).to(dtype)      # noqa: E731
    return dict(z_main=r(b, c, t, h, w), z_edit=r(b, c, k, h, w), v_posi=r(b, c, t + k, h, w), v_nega=r(b, c, t + k, h, w))


# ---------------------------------------------------------------------------------------------------------------
# VAE tiling layer (wan_video_vae.py:1081-1248) around a stand-in convolutional model
# ---------------------------------------------------------------------------------------------------------------
class ToyVAEModel:
    """Stand-in for ``VideoVAE_`` (wan_video_vae.py:951-1056) with its shape contract -- decode (1, 16, T, h, w) ->
    (1, 3, 4T - 3, 8h, 8w), encode (1, 3, T, H, W) -> (1, 16, (T + 3) // 4, H / 8, W / 8) -- cheap, deterministic, and
    dependent on the position INSIDE the tile, so that overlapping tiles disagree and the blending matters.  The tiling
    layer under test never looks inside the model; the real one cannot travel to the GPU box."""

    def decode(self, z, scale):
        x = 1.5 * (z[:, 0:3] * 0.6 + z[:, 3:6] * 0.3 - z[:, 6:9] * 0.2)
        x = x.repeat_interleave(4, dim=2)[:, :, 3:].repeat_interleave(8, dim=3).repeat_interleave(8, dim=4)
        hh, ww = x.shape[3], x.shape[4]
        ry = torch.linspace(0, 1, hh, device=x.device).view(1, 1, 1, hh, 1)
        rx = torch.linspace(0, 1, ww, device=x.device).view(1, 1, 1, 1, ww)
        return (x + (0.2 * ry - 0.1 * rx).to(x.dtype)).to(z.dtype)

    def encode(self, x, scale):
        y = torch.nn.functional.avg_pool3d(x[:, :, ::4].float(), (1, 8, 8)).to(x.dtype)          # (1, 3, (T+3)//4, H/8, W/8)
        mix = torch.stack([y[:, i % 3] * (0.3 + 0.1 * i) - y[:, (i + 1) % 3] * 0.05 * i for i in range(16)], dim=1)
        hh, ww = mix.shape[3], mix.shape[4]
        ry = torch.linspace(0, 1, hh, device=x.device).view(1, 1, 1, hh, 1)
        rx = torch.linspace(0, 1, ww, device=x.device).view(1, 1, 1, 1, ww)
        return (mix + (0.1 * ry + 0.2 * rx).to(mix.dtype)).to(x.dtype)


def vae_build_1d_mask(length, left_bound, right_bound, border_width):
    """WanVideoVAE.build_1d_mask (:1081-1087)."""
    x = torch.ones((length,))
    if not left_bound:
        x[:border_width] = (torch.arange(border_width) + 1) / border_width
    if not right_bound:
        x[-border_width:] = torch.flip((torch.arange(border_width) + 1) / border_width, dims=(0,))
    return x


def vae_build_mask(data, is_bound, border_width):
    """WanVideoVAE.build_mask (:1090-1100)."""
    hh, ww = data.shape[3], data.shape[4]
    h = vae_build_1d_mask(hh, is_bound[0], is_bound[1], border_width[0]).view(hh, 1).expand(hh, ww)
    w = vae_build_1d_mask(ww, is_bound[2], is_bound[3], border_width[1]).view(1, ww).expand(hh, ww)
    return torch.stack([h, w]).min(dim=0).values.view(1, 1, 1, hh, ww)


def vae_tiled(model, source, tile_size, tile_stride, mode, z_dim=16, factor=8, scale=None):
    """WanVideoVAE.tiled_decode (:1103-1153) / tiled_encode (:1155-1204) on ``source``'s own device."""
    _, _, t, hgt, wid = source.shape
    (size_h, size_w), (stride_h, stride_w) = tile_size, tile_stride
    tasks = []
    for h in range(0, hgt, stride_h):
        if h - stride_h >= 0 and h - stride_h + size_h >= hgt:
            continue
        for w in range(0, wid, stride_w):
            if w - stride_w >= 0 and w - stride_w + size_w >= wid:
                continue
            tasks.append((h, h + size_h, w, w + size_w))
    if mode == "decode":
        out_t, oh, ow, ch = t * 4 - 3, hgt * factor, wid * factor, 3
        border = ((size_h - stride_h) * factor, (size_w - stride_w) * factor)
        pos = lambda v: v * factor           # noqa: E731
    else:
        out_t, oh, ow, ch = (t + 3) // 4, hgt // factor, wid // factor, z_dim
        border = ((size_h - stride_h) // factor, (size_w - stride_w) // factor)
        pos = lambda v: v // factor          # noqa: E731
    weight = torch.zeros((1, 1, out_t, oh, ow), dtype=source.dtype, device=source.device)
    values = torch.zeros((1, ch, out_t, oh, ow), dtype=source.dtype, device=source.device)
    for h, h_, w, w_ in tasks:
        tile = source[:, :, :, h:h_, w:w_]
        tile = model.decode(tile, scale) if mode == "decode" else model.encode(tile, scale)
        mask = vae_build_mask(tile, (h == 0, h_ >= hgt, w == 0, w_ >= wid), border).to(dtype=source.dtype, device=source.device)
        th, tw = pos(h), pos(w)
        values[:, :, :, th:th + tile.shape[3], tw:tw + tile.shape[4]] += tile * mask
        weight[:, :, :, th:th + tile.shape[3], tw:tw + tile.shape[4]] += mask
    values = values / weight
    return values.clamp_(-1, 1) if mode == "decode" else values


VAE_CASES = {
    # name: (mode, source shape, tile_size, tile_stride) -- tile sizes in the units the reference's tiled_* methods take
    "decode_small": ("decode", (1, 16, 3, 9, 13), (4, 6), (2, 3)),
    "decode_ragged": ("decode", (1, 16, 2, 7, 10), (4, 4), (3, 2)),
    "decode_one_tile": ("decode", (1, 16, 2, 4, 4), (34, 34), (18, 16)),
    "encode_small": ("encode", (1, 3, 9, 72, 104), (32, 48), (16, 24)),
    "encode_odd": ("encode", (1, 3, 5, 56, 88), (40, 40), (24, 16)),
}


def make_vae_source(shape, seed=11, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g).to(dtype)
