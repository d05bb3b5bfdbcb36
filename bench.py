#!/usr/bin/env python
"""bench.py -- MARCONet inference hot path on B200 (metric: SR chars/s, 128-px-high output).

A "step" is one pass of the full hot path (TextContextEncoderV2 -> TSPGAN -> TSPSRNet, the data flow of the
reference's test_sr.py:145-197) over one batch of synthetic 32x512 LR text lines with 16 characters each
(BASELINE.json configs[1]; --lines sets lines per GPU per step).  One process per GPU; lines are independent
(test_sr.py:77) so ranks shard lines with no data-path collective ("weak" scaling).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--lines L] [--chars C]

Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for how every field is produced.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# algorithmic work, SURVEY.md section 8d (2*MAC of every conv/linear/matmul of the reference modules)
GF_ENCODER_LINE = 111.692
GF_TSPGAN_CHAR = 41.785
GF_SR_LINE = 484.146
GF_SR_CHAR = 47.245


def gflop_per_line(chars):
    return GF_ENCODER_LINE + GF_SR_LINE + (GF_TSPGAN_CHAR + GF_SR_CHAR) * chars


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled DURING the timed region, in-process through NVML (pynvml), every 20 ms;
    falls back to polling nvidia-smi when NVML is unavailable."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def run(self):
        n = self.nvml
        while not self.stop_flag.is_set():
            try:
                if n is not None:
                    mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                    r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                    flags = [bool(r & n.nvmlClocksEventReasonHwSlowdown), bool(r & n.nvmlClocksEventReasonHwThermalSlowdown),
                             bool(r & n.nvmlClocksEventReasonSwThermalSlowdown), bool(r & n.nvmlClocksEventReasonSwPowerCap)]
                    self.samples.append([str(mhz), str(self.max_mhz)] + ["Active" if f else "Not Active" for f in flags])
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.samples.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.02 if n is not None else 0.2)

    def summary(self):
        self.stop_flag.set()
        sm, mx, reasons = [], 0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx = max(mx, float(s[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=mx or None, reasons=sorted(reasons), samples=len(sm),
                    source="nvml" if self.nvml is not None else "nvidia-smi")


def make_inputs(lines, chars, seed):
    from marconet_b200.testing import synth   # seeded synthetic input generators (random tensors only)
    lq = synth.make_lq(lines, seed)
    labels = [synth.make_labels(chars, seed + b) for b in range(lines)]
    locs = synth.make_locs(lines, chars)
    return lq, labels, locs


# --------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port (the reference is Python and cannot travel to the GPU box;
# oracle/restate.py is bit-identical to its modules, tests/test_oracle.py) on all host cores.
# --------------------------------------------------------------------------------------------
_BEST_THREADS = None


def best_cpu_threads():
    """The reference's torch CPU path does not scale to every core of a big host (grouped conv, networks.py:294):
    probe a 1-character TSPGAN forward at a few thread counts and give the baseline its fastest setting."""
    global _BEST_THREADS
    if _BEST_THREADS is None:
        import torch
        from oracle import restate, synth
        sds = synth.make_checkpoints(0)
        ncpu = os.cpu_count() or 1
        cands = sorted({c for c in (ncpu, ncpu // 2, 64, 32, 16, 8) if 1 <= c <= ncpu}, reverse=True)
        lab, sty = synth.make_labels(1, 0), synth.make_styles(1, 0)
        best = None
        with torch.no_grad():
            for c in cands:
                torch.set_num_threads(c)
                restate.tspgan_forward(sds["tspgan"], sty, lab)
                t0 = time.perf_counter()
                restate.tspgan_forward(sds["tspgan"], sty, lab)
                dt = time.perf_counter() - t0
                if best is None or dt < best[0]:
                    best = (dt, c)
        _BEST_THREADS = best[1]
    return _BEST_THREADS


def cpu_line_seconds(chars, repeats=1, threads=None):
    import torch
    from oracle import restate, synth
    threads = threads or best_cpu_threads()
    torch.set_num_threads(threads)
    sds = synth.make_checkpoints(0)
    lq, labels, locs = make_inputs(1, chars, 0)
    best = None
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            restate.full_line(sds, lq, labels, locs)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return best, threads


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    chars = args.chars
    budget_s = 150.0
    t_first, threads = cpu_line_seconds(chars, 1)                      # warm-up step (also sizes the run)
    steps = max(1, min(args.steps, int(budget_s // max(t_first, 1e-3))))
    times = []
    for _ in range(steps):
        t, _ = cpu_line_seconds(chars, 1)
        times.append(t)
    ms = 1e3 * sum(times) / len(times)
    value = chars / (ms / 1e3)
    rec = {
        "impl": "reference", "metric": "sr_chars_per_sec", "value": value, "unit": "chars/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"1 synthetic 32x512 LR line x {chars} chars, encoder->TSPGAN->TSPSRNet (BASELINE configs[1])",
                   "lines_per_step": 1, "chars_per_line": chars},
        "cpu_baseline": {"value": value, "unit": "chars/s", "cores": threads, "kind": "port",
                         "sample": f"{steps} x one full {chars}-char line through oracle/restate.py (bit-identical restatement of the "
                                   f"reference's torch CPU path), torch {torch.__version__}, {threads} threads"},
        "e2e": {"value": value, "unit": "chars/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(rec), flush=True)


# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from marconet_b200 import _lib, ops
    from marconet_b200.models import networks
    from marconet_b200.testing import synth   # synthetic checkpoint generator (random weights only)

    _lib.load()   # fail loudly if the CUDA library is missing
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.precision is not None:
        ops.set_default_precision(args.precision)

    sds = synth.make_checkpoints(0)
    nets = {}
    for key, cls in (("tspgan", networks.TSPGAN), ("encoder", networks.TextContextEncoderV2), ("sr", networks.TSPSRNet)):
        m = cls()
        m.load_state_dict(sds[key], strict=True)
        nets[key] = m.eval().to(dev)

    lines, chars = args.lines, args.chars
    lq_h, labels, locs_h = make_inputs(lines, chars, seed=100 * rank)
    lq_pin, locs_pin = lq_h.pin_memory(), locs_h.pin_memory()
    lab_dev = [l.to(dev) for l in labels]
    lq_dev, locs_dev = lq_pin.to(dev), locs_pin.to(dev)

    lab_all = torch.cat(lab_dev, dim=0)

    def step(lq, locs):
        _, _, w = nets["encoder"](lq)
        # one generator call for the characters of all lines of the step (per-character style = its line's w), then per-line views
        _, f64, f32_ = nets["tspgan"](styles=w.repeat_interleave(chars, dim=0), labels=lab_all, noise=None)
        p64 = [f64[b * chars:(b + 1) * chars] for b in range(lines)]
        p32 = [f32_[b * chars:(b + 1) * chars] for b in range(lines)]
        return nets["sr"](lq, p64, p32, locs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            step(lq_dev, locs_dev)
        if args.profile:   # exactly one step between cudaProfilerStart/Stop (ncu --profile-from-start off)
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            step(lq_dev, locs_dev)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
            return
        sampler = ClockSampler(local)
        sampler.start()
        l0 = ops.LAUNCHES
        ms_total = timed(lambda: step(lq_dev, locs_dev), args.steps)
        launches = (ops.LAUNCHES - l0) // args.steps
        # the same step recorded once into a CUDA graph and replayed: identical kernels and results, no Python
        # between launches, label/window checks on the device.  Used for `value` when capture works and replays bit-exactly.
        graph_info, g = None, None
        if not args.no_graph:
            try:
                from marconet_b200.graph import GraphedLines
                g = GraphedLines(nets["encoder"], nets["tspgan"], nets["sr"], lines=lines, chars=chars, device=dev)
                g.load(lq_dev, lab_all, locs_dev)
                same = bool(torch.equal(g.replay(), step(lq_dev, locs_dev)))
                g.check()
                graph_info = {"launches_per_replay": int(g.launches), "bit_identical_to_eager": same}
            except Exception as exc:   # capture is an optimisation: report why it was skipped and keep the eager number
                graph_info, g = {"error": f"{type(exc).__name__}: {exc}"[:300]}, None
            ok = torch.tensor([1.0 if (g is not None and graph_info["bit_identical_to_eager"]) else 0.0], device=dev)
            if world > 1:              # timed() contains collectives: every rank replays, or none does
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() > 0.5:
                for _ in range(2):
                    g.replay()
                graph_info["ms_per_step"] = timed(g.replay, args.steps) / args.steps
            elif g is not None:
                graph_info["skipped"] = "another rank could not capture" if graph_info["bit_identical_to_eager"] else "replay differs from eager"

        clocks = sampler.summary()      # sampled over the eager and the graph timed regions

        # e2e: host buffers in, SR image out, copies inside the timed region
        sr_host = torch.empty((lines, 3, 128, 2048), dtype=torch.float32).pin_memory()

        def e2e_step():
            lq = lq_pin.to(dev, non_blocking=True)
            locs = locs_pin.to(dev, non_blocking=True)
            sr = step(lq, locs)
            sr_host.copy_(sr, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        e2e_step()
        ms_e2e = timed(e2e_step, args.steps)

        roof = modconv_roofline(nets["tspgan"], chars, dev) if rank == 0 else None

        # informational: end to end through GraphedLines (this repo's own extension API) with the same host buffers and copies.
        # Single process only (no collectives inside, so a failure here cannot desynchronise ranks); runs last.
        if world == 1 and g is not None and graph_info and "ms_per_step" in graph_info:
            try:
                def e2e_graph_step():
                    out = g(lq_pin, None, locs_pin)            # H2D into the static buffers + one graph replay
                    sr_host.copy_(out, non_blocking=True)
                    torch.cuda.current_stream().synchronize()

                e2e_graph_step()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                ev0.record()
                for _ in range(args.steps):
                    e2e_graph_step()
                ev1.record()
                torch.cuda.synchronize()
                g.check()
                graph_info["e2e_ms_per_step"] = ev0.elapsed_time(ev1) / args.steps
                graph_info["e2e_chars_per_sec"] = lines * chars / (graph_info["e2e_ms_per_step"] / 1e3)
            except Exception as exc:
                graph_info["e2e_error"] = f"{type(exc).__name__}: {exc}"[:200]

    total_chars = world * lines * chars
    ms_step = ms_total / args.steps
    ms_eager = ms_step
    # ms_per_step of both modes is already the max over ranks, so every rank takes the same branch
    use_graph = bool(graph_info and "ms_per_step" in graph_info and graph_info["ms_per_step"] < ms_step)
    if use_graph:
        ms_step = graph_info["ms_per_step"]
        launches = graph_info["launches_per_replay"]
    value = total_chars / (ms_step / 1e3)
    e2e_value = total_chars / (ms_e2e / args.steps / 1e3)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            t, threads = cpu_line_seconds(chars, 1)
            cpu = {"value": chars / t, "unit": "chars/s", "cores": threads, "kind": "port",
                   "sample": f"one full {chars}-char 32x512 line through oracle/restate.py (torch CPU fp32, {threads} threads)"}
        rec = {
            "metric": "sr_chars_per_sec", "value": value, "unit": "chars/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{lines} synthetic 32x512 LR line(s) x {chars} chars per GPU per step, "
                                   f"encoder->TSPGAN->TSPSRNet (BASELINE configs[1])",
                       "lines_per_step_per_gpu": lines, "chars_per_line": chars, "parallelism": f"line-sharded dp{world}, no collective",
                       "precision": {0: "fp32 CUDA-core", 1: "fp16x3 tcgen05", 2: "bf16x3 tcgen05", 3: "fp16 tcgen05"}[ops.default_precision()],
                       "l2": "weights (352 MB fp32) + activations (>1 GB/line) exceed the 126 MB L2; no flush needed",
                       "launch_mode": "cuda_graph_replay" if use_graph else "eager_module_calls", "eager_ms_per_step": ms_eager,
                       "cuda_graph": graph_info,
                       "gflop_per_step_per_gpu": gflop_per_line(chars) * lines,
                       "achieved_tflops_per_gpu": gflop_per_line(chars) * lines / ms_step},
            "ms_per_line": ms_step / lines,
            "e2e": {"value": e2e_value, "unit": "chars/s", "h2d_bytes_per_step": int(lq_pin.numel() * 4 + locs_pin.numel() * 4),
                    "d2h_bytes_per_step": int(sr_host.numel() * 4)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        print(json.dumps(rec), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _time_modconv(e, chars, dev, iters):
    import torch
    from marconet_b200 import ops
    x = torch.randn(chars, 32, 32, 512, device=dev)
    dm = torch.rand(chars, 512, device=dev) + 0.5
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)   # 256 MB > L2
    st = torch.cuda.current_stream()
    tot = 0.0
    for i in range(iters + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        ops.conv2d(x, e["w"], 3, 3, pad=(1, 1), bias=e["bias"], out_scale=dm, act=ops.ACT_LRELU02, gain=2 ** 0.5)
        e1.record(st)
        torch.cuda.synchronize()
        if i >= 3:
            tot += e0.elapsed_time(e1)
    ms = tot / iters
    return ms, 2.0 * 512 * 512 * 9 * 32 * 32 * chars


def modconv_roofline(tspgan, chars, dev, iters=20):
    """Live CUDA-event timing of the dominant kernel: the 3x3 modulated conv 512->512 at 32x32 for `chars`
    characters (reference networks.py:294,299 grouped conv; 4.83 GFLOP per character and launch)."""
    from marconet_b200 import ops
    peaks = load_peaks()
    gen = tspgan.TextGenerator
    pk = gen._get_packed(dev)
    e = pk["styled"][6]                                   # convs.5: 512->512 @ 32x32, no upsample
    ms, flops = _time_modconv(e, chars, dev, iters)
    achieved = flops / (ms * 1e-3) / 1e12
    ms_big, flops_big = _time_modconv(e, 128, dev, 5)     # same kernel with 128 characters: 27.7 waves instead of 3.46
    big = flops_big / (ms_big * 1e-3) / 1e12
    passes = {0: 0, 1: 3, 2: 3, 3: 1}[ops.default_precision()]
    return {"kernel": "mn_conv2d_nhwc modulated 3x3 512->512 @32x32 x%d chars (conv_tc2_kernel<128>)" % chars, "bound": "tensor",
            "achieved": achieved, "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": achieved / peaks["tf_burst"],
            # dram__bytes_read.sum + dram__bytes_write.sum of this launch from profiles/r1_tc2_ncu_full.txt (ncu --set full)
            "traffic": 46.6e6 if chars == 16 else None,
            "algorithmic_gflop_per_launch": flops / 1e9, "ms_per_launch": ms,
            "mma_passes_per_algorithmic_flop": passes,
            "tensor_pipe_frac": (achieved * passes / peaks["tf_burst"]) if passes else 0.0,
            "same_kernel_128_chars": {"achieved": big, "frac": big / peaks["tf_burst"], "ms_per_launch": ms_big,
                                      "tensor_pipe_frac": (big * passes / peaks["tf_burst"]) if passes else 0.0},
            "peak_source": peaks["source"] + ", bf16 dense burst (kernel timed alone, L2 flushed between launches); "
                           "fp32-grade products cost 3 fp16 MMAs each, so the ceiling of this kernel is peak/3 in algorithmic terms"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lines", type=int, default=1, help="LR lines per GPU per step")
    ap.add_argument("--chars", type=int, default=16)
    ap.add_argument("--precision", type=int, default=None, help="0 fp32 CUDA-core, 1 fp16x3 tcgen05 (default), 2 bf16x3, 3 fp16x1")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="skip the CUDA-graph replay measurement (value = eager module calls)")
    ap.add_argument("--profile", action="store_true", help="run one step inside cudaProfilerStart/Stop and exit (for ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
