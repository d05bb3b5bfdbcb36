"""GPU LIBRARY baseline (BASELINE.md 5.5 / SURVEY 2.1: "the same reference modules on cuda through stock PyTorch -- cuDNN grouped
conv, cuBLAS, ATen -- is the kernel-level bar to beat on the same box").

The reference itself cannot travel to the GPU box; oracle/restate.py is its bit-identical functional-torch restatement
(tests/test_oracle.py::test_oracle_is_bit_identical_to_reference_modules), so its state dicts are moved to cuda:0 and the same
16-character 32x512 line (BASELINE configs[1]) is timed through stock PyTorch: fp32, TF32 OFF (the reference's arithmetic), with
and without cudnn.benchmark autotuning.  A script, not a pytest test (it lives under tests/ because it executes oracle/):

    python tests/gpu_torch_baseline.py [--chars 16] [--steps 10]

Prints one JSON line; the product path is timed next to it by bench.py on the same box.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chars", type=int, default=16)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--compare", action="store_true", help="also run the product path and report max-abs differences")
    args = ap.parse_args()
    from oracle import restate, synth
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda:0")
    sds = synth.make_checkpoints(0)
    sds_dev = {k: {n: t.to(dev) for n, t in sd.items()} for k, sd in sds.items()}
    lq = synth.make_lq(1, 0).to(dev)
    labels = [synth.make_labels(args.chars, 0).to(dev)]
    locs = synth.make_locs(1, args.chars).to(dev)

    def stage_times(bench_flag):
        torch.backends.cudnn.benchmark = bench_flag
        res = {}
        with torch.no_grad():
            def enc():
                return restate.encoder_forward(sds_dev["encoder"], lq)

            _, _, w = enc()

            def gen():
                return restate.tspgan_forward(sds_dev["tspgan"], w.repeat(args.chars, 1), labels[0])

            _, f64, f32_ = gen()

            def sr():
                return restate.tspsr_forward(sds_dev["sr"], lq, [f64], [f32_], locs)

            def line():
                return restate.full_line(sds_dev, lq, labels, locs)

            for name, fn in (("encoder", enc), ("tspgan", gen), ("tspsr", sr), ("line", line)):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.steps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                res[name + "_ms"] = e0.elapsed_time(e1) / args.steps
        return res

    rec = {"what": "stock PyTorch on cuda (cuDNN/cuBLAS/ATen, fp32, TF32 off) running the reference's op sequence (oracle/restate.py), "
                   f"1 line x {args.chars} chars", "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
           "cudnn_benchmark_off": stage_times(False), "cudnn_benchmark_on": stage_times(True)}
    best = min(rec["cudnn_benchmark_off"]["line_ms"], rec["cudnn_benchmark_on"]["line_ms"])
    rec["best_ms_per_line"] = best
    rec["best_chars_per_sec"] = args.chars / (best / 1e3)
    if args.compare:
        from marconet_b200.models import networks
        nets = {}
        for key, cls in (("tspgan", networks.TSPGAN), ("encoder", networks.TextContextEncoderV2), ("sr", networks.TSPSRNet)):
            m = cls()
            m.load_state_dict(sds[key], strict=True)
            nets[key] = m.eval().to(dev)
        with torch.no_grad():
            ref = restate.full_line(sds_dev, lq, labels, locs)
            _, _, w = nets["encoder"](lq)
            img, f64, f32_ = nets["tspgan"](styles=w.repeat(args.chars, 1), labels=labels[0], noise=None)
            out = nets["sr"](lq, [f64], [f32_], locs)
        rec["product_vs_torch_cuda_max_abs"] = {"sr": float((out - ref["sr"]).abs().max()), "prior": float((img - ref["prior"][0]).abs().max()),
                                                "fea64": float((f64 - ref["fea64"][0]).abs().max())}
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
