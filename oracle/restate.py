"""CPU restatement of the MARCONet inference path (TEST INFRASTRUCTURE ONLY).

Every function takes the *reference-format* ``state_dict`` (same keys/shapes
the reference checkpoints carry) and plain tensors, and
follows the reference arithmetic op for op so that fp32 results agree with the
reference modules to rounding noise.  No nn.Module, no CUDA, no dependency on
``marconet_b200``.  ``dtype=torch.float64`` gives the tie-breaker reference.

Reference files restated (paths relative to the reference repo):
  models/networks.py:27-533, models/resnet.py:1-74, models/textvit_arch.py:1-181
Third-party op restated: basicsr.ops.fused_act (not vendored by the reference,
no version pin; published formula ``leaky_relu(x + b, 0.2) * sqrt(2)``), call
sites models/networks.py:10,195,241.
"""
import math

import torch
import torch.nn.functional as F

SQRT2 = 2 ** 0.5


# --------------------------------------------------------------------------
# third-party: basicsr.ops.fused_act  (call sites models/networks.py:195,241,245)
# --------------------------------------------------------------------------
def fused_leaky_relu(x, bias=None, negative_slope=0.2, scale=SQRT2):
    if bias is not None:
        x = x + bias.view(1, -1, *([1] * (x.dim() - 2)))
    return F.leaky_relu(x, negative_slope) * scale


def _cast(sd, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


# --------------------------------------------------------------------------
# Encoder: ResNet-45 (models/resnet.py:11-74)
# --------------------------------------------------------------------------
_RESNET_BLOCKS = (3, 4, 6, 6, 3)
_RESNET_STRIDES = ((2, 1), (1, 1), (2, 1), (1, 1), (1, 1))


def resnet45(sd, x, prefix="resnet."):
    """models/resnet.py:63-71 (ResNet.forward) + :21-30 (BasicBlock.forward)."""
    x = F.relu(F.conv2d(x, sd[prefix + "conv1.weight"], stride=1, padding=1))
    for li, (nblk, stride) in enumerate(zip(_RESNET_BLOCKS, _RESNET_STRIDES), 1):
        for bi in range(nblk):
            p = f"{prefix}layer{li}.{bi}."
            s = stride if bi == 0 else (1, 1)
            out = F.relu(F.conv2d(x, sd[p + "conv1.weight"]))
            out = F.conv2d(out, sd[p + "conv2.weight"], stride=s, padding=1)
            if (p + "downsample.0.weight") in sd:
                res = F.conv2d(x, sd[p + "downsample.0.weight"], stride=s)
            else:
                res = x
            x = F.relu(out + res)
    return x


# --------------------------------------------------------------------------
# Encoder: TextViT (models/textvit_arch.py:12-181)
# --------------------------------------------------------------------------
def posemb_sincos_2d(h, w, dim, dtype, temperature=10000):
    """models/textvit_arch.py:170-181."""
    y, x = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    omega = torch.arange(dim // 4) / (dim // 4 - 1)
    omega = 1.0 / (temperature ** omega)
    y = y.flatten()[:, None] * omega[None, :]
    x = x.flatten()[:, None] * omega[None, :]
    pe = torch.cat((x.sin(), x.cos(), y.sin(), y.cos()), dim=1)
    return pe.to(dtype)


def _attention(sd, p, x, heads=8):
    """models/textvit_arch.py:104-112."""
    b, n, _ = x.shape
    xn = F.layer_norm(x, (x.shape[-1],), sd[p + "norm.weight"], sd[p + "norm.bias"])
    qkv = F.linear(xn, sd[p + "to_qkv.weight"]).chunk(3, dim=-1)
    q, k, v = [t.reshape(b, n, heads, -1).permute(0, 2, 1, 3) for t in qkv]
    dots = torch.matmul(q, k.transpose(-1, -2)) * (q.shape[-1] ** -0.5)
    attn = dots.softmax(dim=-1)
    out = torch.matmul(attn, v).permute(0, 2, 1, 3).reshape(b, n, -1)
    return F.linear(out, sd[p + "to_out.weight"])


def _feedforward(sd, p, x):
    """models/textvit_arch.py:81-91."""
    h = F.layer_norm(x, (x.shape[-1],), sd[p + "net.0.weight"], sd[p + "net.0.bias"])
    h = F.gelu(F.linear(h, sd[p + "net.1.weight"], sd[p + "net.1.bias"]))
    return F.linear(h, sd[p + "net.3.weight"], sd[p + "net.3.bias"])


def _block(sd, p, x):
    x = _attention(sd, p + "0.", x) + x
    return _feedforward(sd, p + "1.", x) + x


def textvit(sd, feat, prefix="transformer."):
    """models/textvit_arch.py:65-77 (TextViT.forward) + :146-164 (Transformer.forward)."""
    b, c, hh, ww = feat.shape
    p1 = p2 = 8
    h, w = hh // p1, ww // p2
    # Rearrange 'b c (h p1) (w p2) -> b h w (p1 p2 c)'  (textvit_arch.py:32-35)
    x0 = feat.reshape(b, c, h, p1, w, p2).permute(0, 2, 4, 3, 5, 1).reshape(b, h, w, p1 * p2 * c)
    x0 = F.linear(x0, sd[prefix + "to_patch_embedding.1.weight"], sd[prefix + "to_patch_embedding.1.bias"])
    pe = posemb_sincos_2d(h, w, x0.shape[-1], x0.dtype).to(x0.device)
    x = x0.reshape(b, h * w, -1) + pe
    t = prefix + "transformer."
    for i in range(2):
        x = _block(sd, f"{t}layers.{i}.", x)
    x_cls = _block(sd, t + "layers_cls.0.", x)
    xt = x.permute(0, 2, 1)
    xt = F.layer_norm(xt, (64,), sd[t + "linear_seq_maxlen.0.weight"], sd[t + "linear_seq_maxlen.0.bias"])
    x16 = F.linear(xt, sd[t + "linear_seq_maxlen.1.weight"], sd[t + "linear_seq_maxlen.1.bias"]).permute(0, 2, 1)
    x_loc = _block(sd, t + "layers_locs.0.", x16)
    x_w = _block(sd, t + "layers_w.0.", x)

    q = prefix + "linear_cls."
    out_cls = F.linear(F.layer_norm(x_cls, (512,), sd[q + "0.weight"], sd[q + "0.bias"]),
                       sd[q + "1.weight"], sd[q + "1.bias"])
    q = prefix + "linear_w_maxlen."
    xw = x_w.permute(0, 2, 1)
    xw = F.linear(F.layer_norm(xw, (64,), sd[q + "0.weight"], sd[q + "0.bias"]),
                  sd[q + "1.weight"], sd[q + "1.bias"]).permute(0, 2, 1)
    q = prefix + "linear_w."
    xw = xw.reshape(b, -1)
    out_w = F.linear(F.layer_norm(xw, (512,), sd[q + "0.weight"], sd[q + "0.bias"]),
                     sd[q + "1.weight"], sd[q + "1.bias"])
    q = prefix + "linear_locs."
    hl = F.layer_norm(x_loc, (512,), sd[q + "0.weight"], sd[q + "0.bias"])
    hl = F.gelu(F.linear(hl, sd[q + "1.weight"], sd[q + "1.bias"]))
    out_locs = torch.sigmoid(F.linear(hl, sd[q + "3.weight"], sd[q + "3.bias"]))
    return out_cls, out_locs.reshape(b, -1), out_w.reshape(b, -1)


def encoder_forward(sd, lq, dtype=torch.float32):
    """TextContextEncoderV2.forward, models/networks.py:42-45."""
    sd = _cast(sd, dtype)
    feat = resnet45(sd, lq.to(dtype))
    return textvit(sd, feat)


# --------------------------------------------------------------------------
# TSPGAN (models/networks.py:51-321)
# --------------------------------------------------------------------------
def _equal_linear(x, weight, bias, lr_mul=1.0, activation=None):
    """models/networks.py:188-198."""
    scale = (1 / math.sqrt(weight.shape[1])) * lr_mul
    b = bias * lr_mul if bias is not None else None
    if activation == "fused_lrelu":
        return fused_leaky_relu(F.linear(x, weight * scale), b)
    return F.linear(x, weight * scale, bias=b)


def _modulated_conv(sd, p, x, style, demodulate, upsample):
    """models/networks.py:281-302."""
    weight = sd[p + "weight"]  # [1, Cout, Cin, k, k]
    _, cout, cin, k, _ = weight.shape
    batch, _, height, width = x.shape
    s = _equal_linear(style, sd[p + "modulation.weight"], sd[p + "modulation.bias"])
    s = s.view(batch, 1, cin, 1, 1)
    w = (1 / math.sqrt(cin * k * k)) * weight * s
    if demodulate:
        demod = torch.rsqrt(w.pow(2).sum([2, 3, 4]) + 1e-8)
        w = w * demod.view(batch, cout, 1, 1, 1)
    w = w.view(batch * cout, cin, k, k)
    x = x.reshape(1, batch * cin, height, width)
    if upsample:
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
        out = F.conv2d(x, w, padding=1, groups=batch)
    else:
        out = F.conv2d(x, w, padding=k // 2, groups=batch)
    return out.view(batch, cout, out.shape[-2], out.shape[-1])


def _styled_conv(sd, p, x, style, upsample):
    """models/networks.py:242-246."""
    out = _modulated_conv(sd, p + "conv.", x, style, True, upsample)
    out = out + sd[p + "bias"]
    return fused_leaky_relu(out, sd[p + "activate.bias"])


def _to_rgb(sd, p, x, style, skip, upsample=True):
    """models/networks.py:313-321."""
    out = _modulated_conv(sd, p + "conv.", x, style, False, False) + sd[p + "bias"]
    if skip is not None:
        if upsample:
            skip = F.interpolate(skip, scale_factor=2, mode="bilinear", align_corners=False)
        out = out + skip
    return torch.tanh(out)


def tspgan_forward(sd, styles, labels, dtype=torch.float32, return_all=False):
    """TSPGAN.forward -> TextGenerator.forward, models/networks.py:61-62,134-164."""
    sd = _cast(sd, dtype)
    g = "TextGenerator."
    x = styles.to(dtype)
    # PixelNorm (networks.py:170-171) + 8 EqualLinear(lr_mul=0.01, fused_lrelu) (:83-89)
    x = x * torch.rsqrt(torch.mean(x ** 2, dim=1, keepdim=True) + 1e-8)
    for i in range(1, 9):
        x = _equal_linear(x, sd[f"{g}style_mlp.{i}.weight"], sd[f"{g}style_mlp.{i}.bias"],
                          lr_mul=0.01, activation="fused_lrelu")
    latent = x
    # SelectText (networks.py:205-215): embedding[label] broadcast to 4x4, labels concatenated along W
    emb = sd[g + "input_text.TextEmbeddings"]
    labels = labels.long()
    if int(labels.min()) < 0 or int(labels.max()) >= emb.shape[0]:
        raise IndexError("label out of range")
    b, l = labels.shape
    out = emb[labels.reshape(-1), :, 0, 0].reshape(b, l, -1).permute(0, 2, 1)  # [b, C, l]
    out = out[:, :, None, :, None].expand(b, out.shape[1], 4, l, 4).reshape(b, -1, 4, 4 * l)
    out = _styled_conv(sd, g + "conv1.", out, latent, False)
    skip = _to_rgb(sd, g + "to_rgb1.", out, latent, None, upsample=False)
    taps = {}
    for j in range(5):
        out = _styled_conv(sd, f"{g}convs.{2 * j}.", out, latent, True)
        out = _styled_conv(sd, f"{g}convs.{2 * j + 1}.", out, latent, False)
        skip = _to_rgb(sd, f"{g}to_rgbs.{j}.", out, latent, skip)
        taps[out.shape[-1]] = out                      # the reference keys its taps on the WIDTH (networks.py:153-158)
        taps[("rgb", out.shape[-1])] = skip
    if return_all:
        return skip, taps[64], taps[32], taps
    return skip, taps[64], taps[32]


# --------------------------------------------------------------------------
# TSPSRNet (models/networks.py:328-533)
# --------------------------------------------------------------------------
def _sn_weight(sd, p):
    """torch.nn.utils.spectral_norm eval branch: W / (u . (W_mat v))."""
    w = sd[p + "weight_orig"]
    sigma = torch.dot(sd[p + "weight_u"], torch.mv(w.flatten(1), sd[p + "weight_v"]))
    return w / sigma


def _sn_conv(sd, p, x, stride=1):
    return F.conv2d(x, _sn_weight(sd, p), sd[p + "bias"], stride=stride, padding=1)


def _lrelu(x):
    return F.leaky_relu(x, 0.2)


def _gn(sd, p, x):
    """models/networks.py:487-490 (32 channels per group, eps 1e-6)."""
    return F.group_norm(x, x.shape[1] // 32, sd[p + "weight"], sd[p + "bias"], eps=1e-6)


def _swish(x):
    return x * torch.sigmoid(x)


def _res_text_block(sd, p, x_in):
    """ResTextBlockV2.forward, models/networks.py:506-516."""
    x = _sn_conv(sd, p + "conv1.", _swish(_gn(sd, p + "norm1.", x_in)))
    x = _sn_conv(sd, p + "conv2.", _swish(_gn(sd, p + "norm2.", x)))
    if (p + "conv_out.weight") in sd:
        x_in = F.conv2d(x_in, sd[p + "conv_out.weight"], sd[p + "conv_out.bias"])
    return x + x_in


def _two_conv(sd, p, x):
    return _sn_conv(sd, p + "2.", _lrelu(_sn_conv(sd, p + "0.", x)))


def _adain(prior, lq):
    """models/networks.py:518-533 (unbiased variance + 1e-5)."""
    def ms(f):
        b, c = f.shape[:2]
        var = f.reshape(b, c, -1).var(dim=2) + 1e-5
        return f.reshape(b, c, -1).mean(dim=2).view(b, c, 1, 1), var.sqrt().view(b, c, 1, 1)
    lm, ls = ms(lq)
    pm, ps = ms(prior)
    return (prior - pm) / ps * ls + lm


def char_window(loc_center, width_total, half):
    """Window integers for one character, models/networks.py:426-441 / :460-474.

    ``loc_center`` is a 0-dim fp32 tensor; the product and truncation are done in
    fp32 exactly as ``(locs[b][2*c] * W).int()`` does.  Returns (x1, x2, y1, y2).
    """
    center = int((loc_center.float() * width_total).int())
    x1 = 0 if center < half else center - half
    x2 = width_total if center + half > width_total else center + half
    y1 = half - int(math.trunc((x2 - x1) / 2))
    y2 = y1 + x2 - x1
    return x1, x2, y1, y2


def _fuse_level(sd, lvl, feat, priors, locs, half):
    """Per-character prior fusion loop, models/networks.py:421-449 (32) / :455-482 (64)."""
    res = torch.zeros_like(feat)
    W = feat.shape[-1]
    wins = []
    for b, pr in enumerate(priors):
        for c in range(pr.shape[0]):
            x1, x2, y1, y2 = char_window(locs[b][2 * c], W, half)
            wins.append((b, c, x1, x2, y1, y2))
            cp = pr[c:c + 1, :, :, y1:y2]
            cl = feat[b:b + 1, :, :, x1:x2]
            fuse = _res_text_block(sd, f"conv_{lvl}_fuse.0.", torch.cat((_adain(cp, cl), cl), dim=1))
            scale = _two_conv(sd, f"conv_{lvl}_scale.", fuse)
            shift = _two_conv(sd, f"conv_{lvl}_shift.", fuse)
            res[b, :, :, x1:x2] = feat[b, :, :, x1:x2] * scale[0] + shift[0]
    return feat + res, wins


def _up2(x):
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)


def tspsr_forward(sd, lq, priors64, priors32, locs, dtype=torch.float32, return_all=False):
    """TSPSRNet.forward, models/networks.py:411-485."""
    sd = _cast(sd, dtype)
    lq = lq.to(dtype)
    priors64 = [p.to(dtype) for p in priors64]
    priors32 = [p.to(dtype) for p in priors32]
    f32_ = _lrelu(_sn_conv(sd, "conv_first_32.0.", lq))
    f16_ = _lrelu(_sn_conv(sd, "conv_first_16.0.", f32_, stride=2))
    f8_ = _sn_conv(sd, "conv_first_8.2.", _lrelu(_sn_conv(sd, "conv_first_8.0.", f16_, stride=2)))
    s16 = _two_conv(sd, "conv_body_16.", torch.cat([_up2(f8_), f16_], dim=1))
    s32 = _two_conv(sd, "conv_body_32.", torch.cat([_up2(s16), f32_], dim=1))

    p32_256 = [_two_conv(sd, "conv_32_to256.", p) for p in priors32]
    s32f, wins32 = _fuse_level(sd, 32, s32, p32_256, locs, 16)

    x = _lrelu(_sn_conv(sd, "conv_up.1.", _up2(s32f)))
    x = _res_text_block(sd, "conv_up.3.", x)
    s64 = _sn_conv(sd, "conv_up.4.", x)

    s64f, wins64 = _fuse_level(sd, 64, s64, priors64, locs, 32)

    x = _lrelu(_sn_conv(sd, "conv_final.0.", s64f))
    x = _lrelu(_sn_conv(sd, "conv_final.3.", _up2(x)))
    x = _res_text_block(sd, "conv_final.5.", x)
    out = torch.tanh(_sn_conv(sd, "conv_final.6.", x))
    if return_all:
        return out, dict(s32=s32, s32f=s32f, s64=s64, s64f=s64f, wins32=wins32, wins64=wins64)
    return out


# --------------------------------------------------------------------------
# Caller-side integer post-processing (test_w.py:34-40)
# --------------------------------------------------------------------------
def clear_labels(logits, n_alphabet=6735):
    """argmax + CTC-style de-dup, test_w.py:34-40.  ``logits`` is [T, 6736]."""
    idx = torch.max(logits, 1)[1]
    out = []
    for i in range(idx.shape[0]):
        if not (i > 0 and idx[i - 1] == idx[i]) and idx[i] < n_alphabet:
            out.append(int(idx[i]))
    return out


def full_line(sds, lq, labels, locs, dtype=torch.float32):
    """test_sr.py:145-197 data flow for one batch of lines (labels/locs supplied by the caller)."""
    logits, enc_locs, w = encoder_forward(sds["encoder"], lq, dtype)
    outs = dict(logits=logits, enc_locs=enc_locs, w=w)
    p64, p32, imgs = [], [], []
    for b in range(lq.shape[0]):
        lab = labels[b]
        img, f64, f32_ = tspgan_forward(sds["tspgan"], w[b:b + 1].repeat(lab.shape[0], 1), lab, dtype)
        imgs.append(img); p64.append(f64); p32.append(f32_)
    sr = tspsr_forward(sds["sr"], lq, p64, p32, locs, dtype)
    outs.update(prior=imgs, fea64=p64, fea32=p32, sr=sr)
    return outs
