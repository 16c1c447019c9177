"""CPU stand-in for ``video_styler_b200.ops`` used ONLY by the tests of the host-side logic (engine / pipeline /
Ulysses orchestration) where no GPU exists.  Same function names and in-place/out semantics as the C-ABI wrappers,
implemented with the oracle's torch expressions.  Never imported by the product package."""
import torch
import torch.nn.functional as F

from oracle import wan_oracle as O

EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RES, EPI_BIAS_GATE_RES, EPI_BIAS_MUL, EPI_BIAS_GELU_T5 = 0, 1, 2, 3, 4, 5


def as_2d(x):
    return x[0] if x.dim() == 3 else x


def _ret(y, out):
    if out is None:
        return y
    out.copy_(y)
    return out


def ln_modulate(x, shift=None, scale=None, weight=None, bias=None, eps=1e-6, out=None):
    y = F.layer_norm(x, (x.shape[-1],), weight, bias, eps)
    if shift is not None:
        y = y * (1 + scale) + shift
    return _ret(y, out)


def make_rope_table(freqs, device):
    tab = torch.zeros(3, 1024, 32, 2, dtype=torch.float64)
    for a, f in enumerate(freqs):
        n, w = f.shape
        tab[a, :n, :w, 0] = f.real
        tab[a, :n, :w, 1] = f.imag
    return tab


def rope_table_from_freqs(freqs, device):
    f = freqs.reshape(-1, 64)
    return torch.stack([f.real, f.imag], dim=-1).to(torch.float64)


def qk_rmsnorm_rope(q, k, wq, wk, eps, rope_table=None, grid=(1, 1, 1), token_offset=0, frame_ids=None,
                    q_out=None, k_out=None):
    def one(t, w_, dst):
        y = O.rms_norm(t, w_, eps)
        if rope_table is not None:
            n = t.shape[0]
            gf, gh, gw = grid
            if tuple(grid) == (0, 0, 0):
                cs = rope_table                                   # per-token mode: (n, 64, 2)
            else:
                idx = torch.arange(n) + token_offset
                pf, ph, pw = (idx // (gh * gw)).clamp(max=gf - 1), (idx // gw) % gh, idx % gw   # pad rows of the last shard
                if frame_ids is not None:
                    pf = frame_ids.long()[pf]
                cs = torch.cat([rope_table[0, pf, :22], rope_table[1, ph, :21], rope_table[2, pw, :21]], dim=1)  # (n,64,2)
            fr = torch.complex(cs[..., 0], cs[..., 1]).unsqueeze(1)                                          # (n,1,64)
            heads = t.shape[1] // 128
            yc = torch.view_as_complex(y.to(torch.float64).reshape(n, heads, 64, 2))
            y = torch.view_as_real(yc * fr).flatten(1).to(t.dtype)
        dst.copy_(y)
        return dst
    q_out = one(q, wq, q if q_out is None else q_out)
    if k is not None:
        k_out = one(k, wk, k if k_out is None else k_out)
    return q_out, k_out


def linear(x, weight, bias=None, epilogue=EPI_BIAS, gate=None, residual=None, out=None):
    y = F.linear(x, weight, bias)
    if epilogue == EPI_BIAS_GELU:
        y = F.gelu(y, approximate="tanh")
    elif epilogue == EPI_BIAS_RES:
        y = residual + y
    elif epilogue == EPI_BIAS_GATE_RES:
        y = residual + gate * y
    elif epilogue == EPI_BIAS_MUL:
        y = y * residual
    elif epilogue == EPI_BIAS_GELU_T5:
        from oracle import aux_oracle as A
        y = A.t5_gelu(y)
    return _ret(y, out)


def linear_grouped(x, weights, biases, out, variant=0):
    n = weights[0].shape[0]
    for i, (w, b) in enumerate(zip(weights, biases)):
        out[:, i * n:(i + 1) * n].copy_(F.linear(x, w, b))
    return out


def attention_bias(q, k, v, num_heads, bias=None, key_mask=None, scale=1.0, out=None):
    lq, lk = q.shape[0], k.shape[0]
    qh, kh, vh = (t.view(t.shape[0], num_heads, 64).transpose(0, 1) for t in (q, k, v))
    s = torch.matmul(qh, kh.transpose(1, 2)) * scale
    b = torch.zeros_like(s)
    if bias is not None:
        idx = (torch.arange(lk)[None, :] - torch.arange(lq)[:, None]) + (lq - 1)
        b = b + bias[:, idx]
    if key_mask is not None:
        b = b.masked_fill(key_mask.view(1, 1, lk) == 0, torch.finfo(q.dtype).min)
    p = torch.softmax((s + b).float(), dim=-1).to(q.dtype)
    return _ret(torch.matmul(p, vh).transpose(0, 1).reshape(lq, num_heads * 64), out)


def editor_step(z_main, z_edit, v_posi, v_nega, frame_to_key, key_idx, cfg_scale, dt, alpha, beta, dsigma=0.0, euler=True):
    from oracle import aux_oracle as A
    join = lambda v: v if (v is None or torch.is_tensor(v)) else torch.cat(list(v), dim=2)      # noqa: E731
    keys = [int(i) for i in key_idx]
    vp, vn = join(v_posi), join(v_nega)
    if euler:
        return A.editor_step(z_main, z_edit, vp, vn, keys, cfg_scale, dt, alpha, beta, dsigma)
    v = vp if vn is None else vn + cfg_scale * (vp - vn)
    vm, ve = torch.split(v, [z_main.shape[2], z_edit.shape[2]], dim=2)
    return A.velocity_correction(z_main, z_edit, vm, ve, keys, dt, alpha, beta)


def attention(q, k, v, num_heads, out=None, scale=None):
    y = O.attention(q.unsqueeze(0), k.unsqueeze(0), v.unsqueeze(0), num_heads)[0]
    return _ret(y, out)


def scale_add(x, y, scale, out=None):
    return _ret(x + y * scale, out)


def cfg_euler_step(x, v_posi, v_nega, cfg_scale, dsigma, out=None):
    v = v_posi if v_nega is None else v_nega + cfg_scale * (v_posi - v_nega)
    return _ret(x + v * dsigma, out)


def gate_residual(x, gate, y, out=None):
    return _ret(x + gate * y, out)


def ulysses_pack_qkv(qkv, heads, world, out=None):
    n = qkv.shape[0]
    hl = heads // world
    y = qkv.view(n, 3, world, hl, 128).permute(2, 0, 1, 3, 4).contiguous()      # (P, n, 3, hl, 128)
    return _ret(y, out)


def ulysses_unpack_out(recv, heads, world, out=None):
    p, n, w = recv.shape
    y = recv.permute(1, 0, 2).reshape(n, p * w)
    return _ret(y, out)


def tile_blend(values, weight, tile, y0, x0, is_bound, border_width):
    from oracle import aux_oracle as A
    mask = A.vae_build_mask(tile, is_bound, border_width).to(values.dtype)
    th, tw = tile.shape[3], tile.shape[4]
    values[:, :, :, y0:y0 + th, x0:x0 + tw] += tile * mask
    weight[y0:y0 + th, x0:x0 + tw] += mask[0, 0, 0]


def tile_finalize(values, weight, clamp=None):
    values.copy_(values / weight)
    if clamp is not None:
        values.clamp_(clamp[0], clamp[1])
    return values
