"""Round-2 golden fixtures from the UNMODIFIED reference modules (build container only; TEST INFRASTRUCTURE ONLY).

    python -m oracle.make_golden2 [priors1024] [lines8] [windows] [seeds] [width2]

* priors1024 -- BASELINE configs[2] / SURVEY 8d config 3: 1024 random character labels x 1024 distinct random style vectors
  through the reference TSPGAN (chunks of 32 characters; per-sample results do not depend on the batch), strided samples of
  (image, fea64, fea32) per chunk plus one fp64 sum per character and output.  The GPU tests compare sub-batches of 16 / 128 /
  1024 characters against it: the tensor-core kernel picks tiles and split-K by batch size.
* lines8 -- 8 lines x 16 characters (SURVEY 8d config 4 shape, per-line seeds, jittered / clipped / overlapping boxes):
  encoder -> TSPGAN -> TSPSRNet with B = 8.
* windows -- the window integers of models/networks.py:426-441 / :460-474 recorded FROM THE REFERENCE LOOP ITSELF: a
  sys.settrace hook on the unmodified TSPSRNet.forward frame reads its locals (b, c, x1, x2, y1, y2) on the line that slices the
  prior; nothing is restated.  Stored for the config2 / ragged / lines8 inputs and for adversarial centres.
* seeds -- config-2 outputs for checkpoint seeds 1..3 (round 1 pinned only seed 0).
* width2 -- TSPGAN with labels [N, 2] (two characters per row): the reference picks its feature taps by WIDTH
  (networks.py:153-158), so the taps are the 32x64 and 16x32 maps.
"""
import linecache
import os
import sys

import numpy as np
import torch

from . import ref_harness, synth
from .make_golden import GOLDEN_DIR, STRIDES, case_inputs, sample


# ------------------------------------------------------------------------------------------------ window tracer
class WindowTracer:
    """Records (b, c, x1, x2, y1, y2) from the locals of the reference's TSPSRNet.forward while it runs."""

    def __init__(self, forward_code):
        self.code = forward_code
        self.rows = []          # (level, b, c, x1, x2, y1, y2)

    def _local(self, frame, event, arg):
        if event == "line":
            text = linecache.getline(frame.f_code.co_filename, frame.f_lineno)
            if "char_lq_f =" in text:           # the line after the prior slice: x1, x2, y1, y2 are final here
                lo = frame.f_locals
                level = 32 if "sq_f_32[" in text else 64
                self.rows.append((level, int(lo["b"]), int(lo["c"]), int(lo["x1"]), int(lo["x2"]), int(lo["y1"]), int(lo["y2"])))
        return self._local

    def _global(self, frame, event, arg):
        if event == "call" and frame.f_code is self.code:
            return self._local
        return None

    def __enter__(self):
        sys.settrace(self._global)
        return self

    def __exit__(self, *exc):
        sys.settrace(None)
        return False


def traced_sr(models, lq, p64, p32, locs):
    code = type(models["sr"]).forward.__code__
    with WindowTracer(code) as tr, torch.no_grad():
        sr = models["sr"](lq, p64, p32, locs)
    return sr, np.asarray(tr.rows, dtype=np.int64)


def run_case(models, inp):
    lq, labels, locs = inp["lq"], inp["labels"], inp["locs"]
    with torch.no_grad():
        logits, enc_locs, w = models["encoder"](lq)
        imgs, p64, p32 = [], [], []
        for b in range(lq.shape[0]):
            img, f64, f32_ = models["tspgan"](styles=w[b:b + 1].repeat(labels[b].shape[0], 1), labels=labels[b], noise=None)
            imgs.append(img); p64.append(f64); p32.append(f32_)
    sr, wins = traced_sr(models, lq, p64, p32, locs)
    return dict(logits=logits, enc_locs=enc_locs, w=w, image=torch.cat(imgs), fea64=torch.cat(p64), fea32=torch.cat(p32), sr=sr,
                windows_traced=wins)


def lines8_inputs():
    locs = synth.make_locs(8, 16, ragged=True, seed=21)
    return dict(lq=synth.make_lq(8, 40), labels=[synth.make_labels(16, 50 + b) for b in range(8)], locs=locs)


# ------------------------------------------------------------------------------------------------ jobs
def job_windows(models):
    """Traced window integers for the round-1 cases (their SR outputs are already pinned) and adversarial centres."""
    out = {}
    for name in ("config2", "ragged"):
        r = run_case(models, case_inputs(name))
        out[name] = r["windows_traced"]
        print(name, "traced windows", r["windows_traced"].shape)
    # adversarial centres (zero priors; only the integers matter)
    g = torch.Generator().manual_seed(77)
    centres = torch.cat([torch.tensor([0.0, 1e-7, 15.999 / 512, 16.0 / 512, 16.001 / 512, 31.5 / 1024, 32.0 / 1024, 495.999 / 512, 496.0 / 512,
                                        496.001 / 512, 511.0 / 512, 0.99999994, 0.5, 0.25 + 2 ** -20]), torch.rand(50, generator=g) * 0.999])
    centres = centres.float()
    n = centres.numel()
    lq = synth.make_lq(1, 3)                      # all characters on ONE line: the integers depend on locs only
    locs = torch.zeros(1, 2 * n)
    locs[0, 0::2] = centres
    locs[0, 1::2] = 14.0 / 512
    p64 = [torch.zeros(n, 256, 64, 64)]
    p32 = [torch.zeros(n, 512, 32, 32)]
    torch.manual_seed(0)
    _, wins = traced_sr(models, lq, p64, p32, locs)
    out["adversarial"] = wins
    out["adversarial_centres"] = centres.numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "windows_traced.npz"), **out)
    print("windows_traced.npz", {k: v.shape for k, v in out.items()})


def job_lines8(models):
    inp = lines8_inputs()
    out = run_case(models, inp)
    rec = dict(logits=sample(out["logits"], STRIDES["logits"]), locs=sample(out["enc_locs"], 1), w=sample(out["w"], 1),
               image=sample(out["image"], STRIDES["image"]), fea64=sample(out["fea64"], STRIDES["fea64"]),
               fea32=sample(out["fea32"], STRIDES["fea32"]), sr=sample(out["sr"], STRIDES["sr"]),
               argmax=out["logits"].argmax(-1).numpy().astype(np.int64), windows_traced=out["windows_traced"],
               sum_sr=np.float64(out["sr"].double().sum().item()))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "lines8.npz"), **rec)
    print("lines8.npz", {k: (v.shape if hasattr(v, "shape") else v) for k, v in rec.items()})


PRIORS_N, PRIORS_CHUNK = 1024, 32
PRIORS_STRIDES = dict(image=1009, fea64=16411, fea32=8209)


def job_priors1024(models):
    labels, styles = synth.make_labels(PRIORS_N, 11), synth.make_styles(PRIORS_N, 11)
    rec = {k: [] for k in ("image", "fea64", "fea32")}
    sums = {k: [] for k in ("image", "fea64", "fea32")}
    for c0 in range(0, PRIORS_N, PRIORS_CHUNK):
        with torch.no_grad():
            outs = models["tspgan"](styles=styles[c0:c0 + PRIORS_CHUNK], labels=labels[c0:c0 + PRIORS_CHUNK], noise=None)
        for k, o in zip(("image", "fea64", "fea32"), outs):
            rec[k].append(sample(o, PRIORS_STRIDES[k]))
            sums[k].append(o.double().flatten(1).sum(1).numpy())
        print("priors1024 chunk", c0, flush=True)
    out = {k: np.stack(v) for k, v in rec.items()}
    out.update({"sum_" + k: np.concatenate(v) for k, v in sums.items()})
    np.savez_compressed(os.path.join(GOLDEN_DIR, "priors1024.npz"), **out)
    print("priors1024.npz", {k: v.shape for k, v in out.items()})


def job_seeds():
    for seed in (1, 2, 3):
        sds = synth.make_checkpoints(seed)
        models = ref_harness.build_reference_models(sds)
        inp = case_inputs("config2")
        out = run_case(models, inp)
        rec = dict(logits=sample(out["logits"], STRIDES["logits"]), locs=sample(out["enc_locs"], 1), w=sample(out["w"], 1),
                   image=sample(out["image"], STRIDES["image"]), fea64=sample(out["fea64"], STRIDES["fea64"]),
                   fea32=sample(out["fea32"], STRIDES["fea32"]), sr=sample(out["sr"], STRIDES["sr"]),
                   argmax=out["logits"].argmax(-1).numpy().astype(np.int64))
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"config2_seed{seed}.npz"), **rec)
        print("seed", seed, "done", flush=True)


def job_width2(models):
    labels = torch.tensor([[5, 6000], [17, 17], [123, 4567]])
    styles = synth.make_styles(3, 5)
    with torch.no_grad():
        img, fa, fb = models["tspgan"](styles=styles, labels=labels, noise=None)
    rec = dict(image=sample(img, 101), tap_a=sample(fa, 257), tap_b=sample(fb, 263), shape_image=np.asarray(img.shape),
               shape_a=np.asarray(fa.shape), shape_b=np.asarray(fb.shape))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "width2.npz"), **rec)
    print("width2.npz", img.shape, fa.shape, fb.shape)


def main():
    jobs = sys.argv[1:] or ["windows", "width2", "lines8", "seeds", "priors1024"]
    sds = synth.make_checkpoints(0)
    models = ref_harness.build_reference_models(sds)
    for j in jobs:
        if j == "seeds":
            job_seeds()
        else:
            globals()["job_" + j](models)


if __name__ == "__main__":
    main()
