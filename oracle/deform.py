"""TEST INFRASTRUCTURE (oracle).  Torch restatement of ``torchvision.ops.deform_conv2d`` (modulated deformable
convolution v2), the third-party operator behind ``DeformConv2d`` at ICIP2023/src/model/m.py:29-34 and
ICIP2024/src/model/helpers.py:40,57.  torchvision 0.26 IS importable in this image, so this restatement is pinned
against the real operator on the CPU (tests/test_oracle_deform.py) and the GPU tests compare the kernel with the real
operator on the device as well.

Algorithm (torchvision/csrc/ops/cuda/deform_conv2d_kernel.cu, deformable_im2col + per-group GEMM):
  for output position (y_o, x_o), kernel point k = (i, j), offset group g (channels c with c // (Cin/OG) == g):
      y = (y_o*stride_h - pad_h) + i*dil_h + offset[n, (g*K + k)*2    , y_o, x_o]
      x = (x_o*stride_w - pad_w) + j*dil_w + offset[n, (g*K + k)*2 + 1, y_o, x_o]
      col[c, k] = mask[n, g*K + k, y_o, x_o] * bilinear(input[n, c], y, x)
  bilinear: 0 if y <= -1 or y >= H or x <= -1 or x >= W; corners outside the image contribute 0
  out[n, co] = bias[co] + sum_{c in weight group of co, k} weight[co, c_local, k] * col[c, k]
"""
import torch


def _bilinear(inp, y, x):
    """inp [N, C, H, W]; y, x [N, 1, P] float -> [N, C, P]."""
    N, C, H, W = inp.shape
    outside = (y <= -1) | (y >= H) | (x <= -1) | (x >= W)
    y0, x0 = torch.floor(y), torch.floor(x)
    lh, lw = y - y0, x - x0
    hh, hw = 1 - lh, 1 - lw
    y0, x0 = y0.long(), x0.long()
    y1, x1 = y0 + 1, x0 + 1
    flat = inp.reshape(N, C, H * W)

    def tap(yy, xx, ok):
        idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).expand(N, C, -1)
        return torch.gather(flat, 2, idx) * ok.to(inp.dtype)

    v1 = tap(y0, x0, (y0 >= 0) & (x0 >= 0))
    v2 = tap(y0, x1, (y0 >= 0) & (x1 <= W - 1))
    v3 = tap(y1, x0, (y1 <= H - 1) & (x0 >= 0))
    v4 = tap(y1, x1, (y1 <= H - 1) & (x1 <= W - 1))
    val = hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4
    return val * (~outside).to(inp.dtype)


def deform_conv2d(input, offset, weight, bias=None, stride=(1, 1), padding=(0, 0), dilation=(1, 1), mask=None):
    pair = lambda v: (v, v) if isinstance(v, int) else tuple(v)
    (sh, sw), (ph, pw), (dh, dw) = pair(stride), pair(padding), pair(dilation)
    N, Cin, H, W = input.shape
    Cout, cin_g, kh, kw = weight.shape
    groups, K = Cin // cin_g, kh * kw
    Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    OG = offset.shape[1] // (2 * K)
    ch_og = Cin // OG
    P = Ho * Wo
    ys = (torch.arange(Ho, device=input.device) * sh - ph).view(Ho, 1).expand(Ho, Wo).reshape(1, 1, P)
    xs = (torch.arange(Wo, device=input.device) * sw - pw).view(1, Wo).expand(Ho, Wo).reshape(1, 1, P)
    cols = input.new_zeros(N, Cin, K, P)
    off = offset.reshape(N, OG, K, 2, P)
    msk = mask.reshape(N, OG, K, P) if mask is not None else None
    for g in range(OG):
        sl = slice(g * ch_og, (g + 1) * ch_og)
        for k in range(K):
            i, j = divmod(k, kw)
            y = (ys + i * dh).to(input.dtype) + off[:, g, k, 0].unsqueeze(1)
            x = (xs + j * dw).to(input.dtype) + off[:, g, k, 1].unsqueeze(1)
            v = _bilinear(input[:, sl], y, x)
            if msk is not None:
                v = v * msk[:, g, k].unsqueeze(1)
            cols[:, sl, k] = v
    cols = cols.reshape(N, groups, cin_g * K, P)
    wmat = weight.reshape(groups, Cout // groups, cin_g * K)
    out = torch.einsum("goc,ngcp->ngop", wmat, cols).reshape(N, Cout, Ho, Wo)
    if bias is not None:
        out = out + bias.view(1, -1, 1, 1)
    return out
