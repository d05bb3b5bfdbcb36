"""TextViT head of the LR encoder (reference models/textvit_arch.py:12-181): parameter container
with the reference's state_dict keys + a fixed kernel sequence.

  patch embedding  = 8x8/stride-8 implicit-GEMM conv over the NHWC ResNet feature map (the
                     reference's Rearrange 'b c (h p1)(w p2) -> b h w (p1 p2 c)' + Linear(32768,512)
                     has exactly the K ordering (ky,kx,c) of that conv), split-K over 148 SMs,
                     sincos positional embedding added in the epilogue;
  attention        = one fused kernel per block (softmax(QK^T/8)V for all heads, S<=64);
  LayerNorm over tokens + Linear over tokens (the permute idiom) = one kernel (mn_token_mix).
"""
import torch
import torch.nn as nn

from .. import ops
from ..ops import ACT_GELU, ACT_SIGMOID


def _lin(m, name=None):
    """nn.Linear -> ([in,out] packed weight, bias or None)."""
    return ops.ConvWeight(m.weight.detach().t().contiguous(), 1, name=name), (None if m.bias is None else m.bias.detach().contiguous())


def _ln(m):
    return m.weight.detach().contiguous(), m.bias.detach().contiguous()


class FeedForward(nn.Module):
    def __init__(self, dim, hidden_dim):
        super().__init__()
        self.net = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, hidden_dim), nn.GELU(), nn.Linear(hidden_dim, dim))


class Attention(nn.Module):
    def __init__(self, dim, heads=8, dim_head=64):
        super().__init__()
        inner = dim_head * heads
        self.heads, self.dim_head, self.scale = heads, dim_head, dim_head ** -0.5
        self.norm = nn.LayerNorm(dim)
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.to_out = nn.Linear(inner, dim, bias=False)


def _block_pack(pair):
    attn, ff = pair
    return dict(ln1=_ln(attn.norm), qkv=_lin(attn.to_qkv)[0], out=_lin(attn.to_out)[0], heads=attn.heads, dh=attn.dim_head,
                ln2=_ln(ff.net[0]), fc1=_lin(ff.net[1]), fc2=_lin(ff.net[3]))


def _block_run(pk, x, b, s):
    """Pre-norm block on x: [b*s, dim] (reference textvit_arch.py:148-150, 104-112, 81-91)."""
    h = ops.layernorm(x, *pk["ln1"])
    qkv = ops.linear(h, pk["qkv"])
    a = ops.attention(qkv.view(b, s, -1), pk["heads"], pk["dh"]).view(b * s, -1)
    x = ops.linear(a, pk["out"], residual=x)
    h = ops.layernorm(x, *pk["ln2"])
    h = ops.linear(h, pk["fc1"][0], pk["fc1"][1], act=ACT_GELU)
    return ops.linear(h, pk["fc2"][0], pk["fc2"][1], residual=x)


class Transformer(nn.Module):
    def __init__(self, dim, depth, heads, dim_head, mlp_dim):
        super().__init__()
        mk = lambda mlp: nn.ModuleList([Attention(dim, heads=heads, dim_head=dim_head), FeedForward(dim, mlp)])
        self.layers = nn.ModuleList([mk(mlp_dim) for _ in range(depth - 1)])
        self.layers_cls = nn.ModuleList([mk(mlp_dim)])
        self.layers_locs = nn.ModuleList([mk(mlp_dim // 2)])
        self.layers_w = nn.ModuleList([mk(mlp_dim // 2)])
        self.linear_seq_maxlen = nn.Sequential(nn.LayerNorm(64), nn.Linear(64, 16))


def posemb_sincos_2d(h, w, dim, device, temperature=10000):
    """Restatement of reference textvit_arch.py:170-181 (host-side constant, computed once per pack)."""
    y, x = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    omega = torch.arange(dim // 4) / (dim // 4 - 1)
    omega = 1.0 / (temperature ** omega)
    y = y.flatten()[:, None] * omega[None, :]
    x = x.flatten()[:, None] * omega[None, :]
    return torch.cat((x.sin(), x.cos(), y.sin(), y.cos()), dim=1).float().to(device)


class TextViT(nn.Module):
    def __init__(self, num_classes, dim, max_length=16):
        super().__init__()
        self.patch, self.dim, self.max_length = 8, dim, 16
        patch_dim = 512 * 8 * 8
        self.to_patch_embedding = nn.Sequential(nn.Identity(), nn.Linear(patch_dim, dim))   # [0] = Rearrange slot
        self.transformer = Transformer(dim, 3, 8, 64, 1024)
        self.to_latent = nn.Identity()
        self.linear_cls = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, num_classes))
        self.linear_locs = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, dim // 2), nn.GELU(), nn.Linear(dim // 2, 2),
                                         nn.Sigmoid())
        self.linear_w = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, 512))
        self.linear_w_maxlen = nn.Sequential(nn.LayerNorm(64), nn.Linear(64, 1))

    def pack(self, name="transformer"):
        t = self.transformer
        dev = self.linear_cls[1].weight.device
        return dict(
            patch=_lin(self.to_patch_embedding[1]), pe=posemb_sincos_2d(1, 64, self.dim, dev).contiguous(),
            layers=[_block_pack(p) for p in t.layers], cls=_block_pack(t.layers_cls[0]),
            locs=_block_pack(t.layers_locs[0]), w=_block_pack(t.layers_w[0]),
            seq=(_ln(t.linear_seq_maxlen[0]), t.linear_seq_maxlen[1].weight.detach().contiguous(),
                 t.linear_seq_maxlen[1].bias.detach().contiguous()),
            head_cls=(_ln(self.linear_cls[0]), _lin(self.linear_cls[1])),
            head_locs=(_ln(self.linear_locs[0]), _lin(self.linear_locs[1]), _lin(self.linear_locs[3])),
            head_w=(_ln(self.linear_w[0]), _lin(self.linear_w[1])),
            wmax=(_ln(self.linear_w_maxlen[0]), self.linear_w_maxlen[1].weight.detach().contiguous(),
                  self.linear_w_maxlen[1].bias.detach().contiguous()),
        )

    def run(self, pk, feat, branch=None):
        """feat: NHWC [B,8,512,512] -> (logits [B,64,6736], locs [B,32], w [B,512]).

        ``branch = (stream, split_k_scratch)``: the classification and box branches -- which the style vector ``w`` does not
        depend on -- are launched on that stream, so that a pipeline that only waits for ``w`` (the prior generator) can go on
        while they run.  The CALLER joins the stream (``main.wait_stream(stream)``) before it reads ``logits`` / ``locs``."""
        b, fh, fw, c = feat.shape
        if fh != 8 or fw % 8 != 0 or c != 512:
            raise RuntimeError(f"TextViT expects a [B,8,W,512] feature map, got {tuple(feat.shape)}")
        s = fw // 8
        if s != 64:
            raise RuntimeError("TextViT is built for 32x512 LR lines (64 tokens)")
        pw, pb = pk["patch"]
        if b <= 4:      # weight-streaming GEMM with M = 64 tokens per line: gathered small-M kernel, K split over a cluster
            x = ops.patch_embed(feat, pw.w, pb, pk["pe"])
        else:
            x = ops.conv2d(feat, pw.w, 8, 8, stride=(8, 8), bias=pb, residual=pk["pe"].view(1, 1, s, -1), res_broadcast=True)
            x = x.view(b * s, self.dim)
        for blk in pk["layers"]:
            x = _block_run(blk, x, b, s)

        def cls_and_locs():
            x_cls = _block_run(pk["cls"], x, b, s)
            (g, be), w16, b16 = pk["seq"]
            x16 = ops.token_mix(x.view(b, s, -1), g, be, w16, b16)                     # [B,16,512]
            x_loc = _block_run(pk["locs"], x16.view(b * 16, -1), b, 16)
            ln, (w, bias) = pk["head_cls"]
            lg = ops.linear(ops.layernorm(x_cls, *ln), w, bias).view(b, s, -1)
            ln, (w1, b1), (w2, b2) = pk["head_locs"]
            hl = ops.linear(ops.layernorm(x_loc, *ln), w1, b1, act=ACT_GELU)
            return lg, ops.linear(hl, w2, b2, act=ACT_SIGMOID).view(b, -1)

        if branch is None:
            logits, locs = cls_and_locs()
        else:
            stream, scratch = branch
            main = torch.cuda.current_stream(feat.device)
            stream.wait_stream(main)
            x.record_stream(stream)
            with torch.cuda.stream(stream), ops.use_workspace(scratch):
                logits, locs = cls_and_locs()
            logits.record_stream(main)
            locs.record_stream(main)

        x_w = _block_run(pk["w"], x, b, s)
        (g, be), w1, b1 = pk["wmax"]
        xw = ops.token_mix(x_w.view(b, s, -1), g, be, w1, b1).view(b, -1)          # [B,512]
        ln, (w, bias) = pk["head_w"]
        out_w = ops.linear(ops.layernorm(xw, *ln), w, bias)
        return logits, locs, out_w

    @torch.no_grad()
    def forward(self, img):
        if not img.is_cuda:
            raise RuntimeError("marconet_b200 TextViT runs only on CUDA (sm_100a) devices")
        return self.run(self.pack(), ops.as_nhwc(img.float()))
