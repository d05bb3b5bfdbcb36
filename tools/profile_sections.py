"""In-situ section timing of one 16-character line (CUDA events on the launching stream, eager module calls, warm caches --
unlike an ncu launch list, which replays every kernel alone with flushed caches).  Developer tool; prints one JSON line.

    python tools/profile_sections.py [--chars 16] [--lines 1] [--iters 20]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chars", type=int, default=16)
    ap.add_argument("--lines", type=int, default=1)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    from marconet_b200 import ops
    from marconet_b200.models import networks
    from marconet_b200.models.networks import _res_block, _two
    from marconet_b200.testing import synth
    dev = torch.device("cuda:0")
    sds = synth.make_checkpoints(0)
    nets = {}
    for key, cls in (("tspgan", networks.TSPGAN), ("encoder", networks.TextContextEncoderV2), ("sr", networks.TSPSRNet)):
        m = cls()
        m.load_state_dict(sds[key], strict=True)
        nets[key] = m.eval().to(dev)
    L, C = args.lines, args.chars
    lq = synth.make_lq(L, 0).to(dev)
    labels = torch.cat([synth.make_labels(C, b) for b in range(L)], 0).to(dev)
    locs = synth.make_locs(L, C).to(dev)
    enc, gen, sr = nets["encoder"], nets["tspgan"], nets["sr"]
    times = {}
    order = []

    class Sec:
        def __init__(self, name):
            self.name = name

        def __enter__(self):
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

        def __exit__(self, *exc):
            self.e1.record()
            if self.name not in times:
                times[self.name] = []
                order.append(self.name)
            times[self.name].append((self.e0, self.e1))
            return False

    def one_pass():
        with torch.no_grad(), Sec("TOTAL"):
            pk = enc._get_packed(dev)
            with Sec("enc.nchw_to_nhwc"):
                x = ops.nchw_to_nhwc(lq)
            with Sec("enc.resnet"):
                feat = enc.resnet.run(pk["resnet"], x)
            with Sec("enc.textvit"):
                logits, locs_lr, w = enc.transformer.run(pk["vit"], feat)
            with Sec("tspgan"):
                img, f64, f32_ = gen(styles=w.repeat_interleave(C, dim=0), labels=labels, noise=None)
            spk = sr._get_packed(dev)
            counts = [C] * L
            with Sec("sr.trunk"):
                s32 = sr._trunk(spk, lq)
            with Sec("sr.to256"):
                p32 = _two(spk["conv_32_to256"], ops.as_nhwc(f32_))
            with Sec("sr.fuse32"):
                s32 = sr._fuse(spk, 32, s32, p32, locs.cpu(), counts, 16)
            with Sec("sr.conv_up"):
                u = ops.resample_modulate(s32, None, up=True)
                x, mr = ops.conv2d(u, spk["up_1"][0], 3, 3, pad=(1, 1), bias=spk["up_1"][1], act=ops.ACT_LRELU02, gn_stats=True)
                x = _res_block(spk["up_res"], x, mr1=mr)
                s64 = ops.conv2d(x, spk["up_4"][0], 3, 3, pad=(1, 1), bias=spk["up_4"][1])
            with Sec("sr.fuse64"):
                s64 = sr._fuse(spk, 64, s64, ops.as_nhwc(f64), locs.cpu(), counts, 32)
            with Sec("sr.conv_final"):
                x = ops.conv2d(s64, spk["fin_0"][0], 3, 3, pad=(1, 1), bias=spk["fin_0"][1], act=ops.ACT_LRELU02)
                u = ops.resample_modulate(x, None, up=True)
                x, mr = ops.conv2d(u, spk["fin_3"][0], 3, 3, pad=(1, 1), bias=spk["fin_3"][1], act=ops.ACT_LRELU02, gn_stats=True)
                x = _res_block(spk["fin_res"], x, mr1=mr)
                out = ops.conv2d(x, spk["fin_6"][0], 3, 3, pad=(1, 1), bias=spk["fin_6"][1], act=ops.ACT_TANH)
        return out

    for _ in range(3):
        one_pass()
    torch.cuda.synchronize()
    times.clear(); order.clear()
    for _ in range(args.iters):
        one_pass()
    torch.cuda.synchronize()
    rec = {"lines": L, "chars": C, "note": "ms per section, mean over iters, eager launches on one stream (host gaps included)"}
    for k in order:
        rec[k] = round(sum(a.elapsed_time(b) for a, b in times[k]) / len(times[k]), 4)
    # the three reference-facing module calls as they run in production: per-signature CUDA-graph replays (no host gaps)
    lab_cpu = labels.cpu()
    with torch.no_grad():
        def t_mod(fn):
            for _ in range(4):
                out = fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                out = fn()
            e1.record()
            torch.cuda.synchronize()
            return round(e0.elapsed_time(e1) / args.iters, 4), out
        rec["module_graph.encoder"], (_, _, w) = t_mod(lambda: enc(lq))
        rec["module_graph.tspgan"], (_, f64, f32_) = t_mod(lambda: gen(styles=w.repeat_interleave(C, dim=0), labels=lab_cpu, noise=None))
        p64 = [f64[b * C:(b + 1) * C] for b in range(L)]
        p32 = [f32_[b * C:(b + 1) * C] for b in range(L)]
        rec["module_graph.tspsr"], _ = t_mod(lambda: sr(lq, p64, p32, locs))
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
