"""PROTOTYPE (host-side, CPU-checkable) of the plan in DESIGN.md section 7 for the Wan VAE's CausalConv3d on the tcgen05
GEMM: the convolution as ONE GEMM with K = taps * C_in whose A operand is `taps` row-shifted windows of a zero-padded
channels-last buffer.  Not product code and not imported by the package: it pins the index arithmetic the TMA producer
would use, against torch's own conv3d, so that the next round starts from a verified formulation.

  input  x (C_in, T, H, W)  ->  padded channels-last buffer P of shape (T + 2, H + 2, W + 2, C_in), flattened to rows:
         causal padding = 2 frames in FRONT of time (wan_video_vae.py:38-52: padding (2p_t, 0) in time), 1 pixel on every
         side in space; row(t, h, w) = (t * (H + 2) + h) * (W + 2) + w
  output computed for EVERY padded position r (row-major over the padded grid): out[r] = sum_tap P[r + off(tap)] @ W_tap^T
         with off(dt, dh, dw) = (dt * (H + 2) + dh) * (W + 2) + dw -- a constant row offset per tap, i.e. per k-block of
         the GEMM: the producer adds it to the TMA row coordinate; rows past the end read as zero (TMA out-of-bounds
         fill).  The rows of interest are r = row(t, h, w) for t < T, h < H, w < W (the window starts at the output
         position), the others are discarded (epilogue mask): (H+2)(W+2) / (H W) of wasted work, 5 % at 60 x 104.
  weight (C_out, C_in, 3, 3, 3) -> (C_out, 27 * C_in), tap-major: the K-major B operand of the GEMM.

    python tools/prototypes/conv3d_shifted_gemm.py        # self-check against F.conv3d, fp32
"""
import torch
import torch.nn.functional as F


def causal_conv3d_reference(x, weight, bias):
    """CausalConv3d.forward without a feature cache (wan_video_vae.py:44-52): pad (1, 1, 1, 1, 2, 0), then conv3d."""
    return F.conv3d(F.pad(x.unsqueeze(0), (1, 1, 1, 1, 2, 0)), weight, bias)[0]


def causal_conv3d_shifted_gemm(x, weight, bias):
    c_in, t, h, w = x.shape
    c_out = weight.shape[0]
    hp, wp = h + 2, w + 2
    padded = torch.zeros(t + 2, hp, wp, c_in, dtype=x.dtype)
    padded[2:, 1:h + 1, 1:w + 1] = x.permute(1, 2, 3, 0)                  # causal: both extra frames in front
    rows = padded.reshape(-1, c_in)                                        # ((T+2)(H+2)(W+2), C_in)
    n_rows = rows.shape[0]
    w_taps = weight.permute(0, 2, 3, 4, 1).reshape(c_out, 27, c_in)       # tap-major K
    out = torch.zeros(n_rows, c_out, dtype=x.dtype)
    for tap in range(27):                                                  # = the k-blocks of ONE GEMM launch
        dt, dh, dw = tap // 9, (tap // 3) % 3, tap % 3
        off = (dt * hp + dh) * wp + dw
        a = torch.zeros_like(rows)                                         # rows past the end read as zero (TMA OOB fill)
        a[:n_rows - off] = rows[off:]
        out += a @ w_taps[:, tap].t()
    out = out + bias
    grid = out.reshape(t + 2, hp, wp, c_out)[:t, :h, :w]                   # the epilogue keeps the rows of real outputs
    return grid.permute(3, 0, 1, 2).contiguous()


if __name__ == "__main__":
    g = torch.Generator().manual_seed(0)
    for (c_in, c_out, t, h, w) in [(4, 6, 3, 5, 7), (8, 8, 1, 4, 4), (16, 12, 5, 9, 6)]:
        x = torch.randn(c_in, t, h, w, generator=g)
        wt = torch.randn(c_out, c_in, 3, 3, 3, generator=g) / (27 * c_in) ** 0.5
        b = torch.randn(c_out, generator=g)
        ref = causal_conv3d_reference(x, wt, b)
        got = causal_conv3d_shifted_gemm(x, wt, b)
        err = float((got - ref).abs().max())
        print(f"C {c_in}->{c_out}, T {t}, {h}x{w}: max |diff| vs F.conv3d = {err:.2e}")
        assert err < 1e-5
    print("ok")
