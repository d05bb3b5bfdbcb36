"""BASELINE configs[4] / SURVEY 8d config 5: the data flow of the reference's test_w.py:95-108 scaled up -- two encoder
passes give w1, w2; labels = the 16 characters of image 1; priors are generated for `--steps` interpolated styles
w = w1*t + w2*(1-t), either one TSPGAN call per step (like the script) or one batched call of steps*16 (char, w) pairs.

    python tools/bench_style_sweep.py [--steps 256] [--chars 16] [--mode per_step|batched] [--chunk 128]

Prints one JSON line: prior characters per second (CUDA events, max of nothing: single GPU), the launch count per TSPGAN call and
the tensor-pipe share implied by the algorithmic 41.785 GFLOP per character.  Not part of the product path.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

GFLOP_PER_CHAR = 41.785      # TSPGAN, algorithmic


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=256, help="interpolation steps (test_w.py uses 11)")
    ap.add_argument("--chars", type=int, default=16)
    ap.add_argument("--mode", default="batched", choices=["per_step", "batched"])
    ap.add_argument("--chunk", type=int, default=128, help="(char, w) pairs per TSPGAN call in batched mode")
    ap.add_argument("--repeat", type=int, default=2)
    args = ap.parse_args()
    from marconet_b200 import ops
    from marconet_b200.models import networks
    from marconet_b200.testing import synth

    dev = torch.device("cuda:0")
    sds = synth.make_checkpoints(0)
    enc, gen = networks.TextContextEncoderV2(), networks.TSPGAN()
    enc.load_state_dict(sds["encoder"], strict=True)
    gen.load_state_dict(sds["tspgan"], strict=True)
    enc, gen = enc.eval().to(dev), gen.eval().to(dev)
    labels = synth.make_labels(args.chars, 7).to(dev)
    with torch.no_grad():
        _, _, w1 = enc(synth.make_lq(1, 21).to(dev))
        _, _, w2 = enc(synth.make_lq(1, 22).to(dev))
        ts = torch.linspace(0, 1, args.steps, device=dev).view(-1, 1)
        w_steps = w1 * ts + w2 * (1 - ts)                                   # [steps, 512]   (test_w.py:101)

        def sweep():
            if args.mode == "per_step":
                for i in range(args.steps):
                    gen(styles=w_steps[i:i + 1].repeat(args.chars, 1), labels=labels, noise=None)
            else:
                styles = w_steps.repeat_interleave(args.chars, dim=0)      # [steps*chars, 512]
                labs = labels.repeat(args.steps, 1)
                for c0 in range(0, styles.shape[0], args.chunk):
                    gen(styles=styles[c0:c0 + args.chunk], labels=labs[c0:c0 + args.chunk], noise=None)

        sweep()
        torch.cuda.synchronize()
        l0 = ops.LAUNCHES
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.repeat):
            sweep()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.repeat
    total = args.steps * args.chars
    calls = args.steps if args.mode == "per_step" else -(-total // args.chunk)
    print(json.dumps({"workload": f"test_w.py sweep: {args.steps} styles x {args.chars} chars, {args.mode}", "prior_chars_per_sec": total / (ms / 1e3),
                      "ms_per_sweep": ms, "tspgan_calls": calls, "launches_per_call": (ops.LAUNCHES - l0) // (args.repeat * calls),
                      "algorithmic_tflops": total * GFLOP_PER_CHAR / ms, "chunk": args.chunk if args.mode == "batched" else args.chars}))


if __name__ == "__main__":
    main()
