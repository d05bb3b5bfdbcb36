#!/usr/bin/env python
"""Summarise an `ncu --set full --import-source on` capture of conv_tc2_kernel (run here, no GPU needed):

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep [--chars 16 --dram-json profiles/r2_tc2_dram.json] > profiles/<name>.txt

Prints the headline metrics (duration, tensor pipe, DRAM bytes, issue slots, XU, shared-memory wavefronts) and the share of
warp-stall samples per kernel region.  Regions are found from landmark SASS instructions of the warp-specialised kernel:
F2FP = operand split, LDTM/STG = epilogue, STTM = TMEM feed, UTCHMMA = MMA issue, UTMALDG = TMA producers; samples on the
mbarrier try-wait spin (TRYWAIT + the ISETP after it) are counted separately as "waiting"."""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__cluster_size", "launch__grid_size", "smsp__inst_executed.sum"]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = ncu_csv(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[-1]
    raw = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print("== headline metrics:", rep)
    for k in KEYS:
        if k in raw:
            print(f"  {k} = {raw[k][0]} {raw[k][1]}")
    if "--dram-json" in sys.argv:
        path = sys.argv[sys.argv.index("--dram-json") + 1]
        chars = int(sys.argv[sys.argv.index("--chars") + 1]) if "--chars" in sys.argv else 16
        mult = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
        tot = sum(float(raw[k][0]) * mult.get(raw[k][1], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        json.dump({"chars": chars, "dram_bytes_per_launch": tot, "capture": "ncu --set full --clock-control none, " + rep.split("/")[-1],
                   "duration_us_under_ncu": float(raw["gpu__time_duration.sum"][0])}, open(path, "w"))
        print("  wrote", path, tot)
    src = ncu_csv(rep, "source")
    h = src[1]
    ix = {n: i for i, n in enumerate(h)}
    data = src[2:]

    def f(r, k):
        try:
            return float(r[ix[k]])
        except Exception:
            return 0.0

    marks = {"split (F2FP)": "F2FP", "epilogue drain (LDTM)": "LDTM", "epilogue store (STG)": "STG.E.128", "TMEM feed (STTM)": "STTM",
             "MMA issue (UTCHMMA)": "UTCHMMA"}
    pos = {k: [i for i, r in enumerate(data) if v in r[ix["Source"]]] for k, v in marks.items()}
    spans = {k: (min(v), max(v)) for k, v in pos.items() if v}
    total = sum(f(r, "# Samples") for r in data)
    print("== warp-stall samples by region (share of all samples; 'work' excludes mbarrier spin-waits)")
    for k, (a, b) in sorted(spans.items(), key=lambda kv: kv[1]):
        seg = data[max(0, a - 60):b + 60]
        wait = sum(f(r, "# Samples") for r in seg if "TRYWAIT" in r[ix["Source"]] or "ISETP.NE.OR" in r[ix["Source"]])
        tot = sum(f(r, "# Samples") for r in seg)
        ex = sum(f(r, "Instructions Executed") for r in seg)
        print(f"  {k:28s} SASS[{a}:{b}] work {100 * (tot - wait) / total:5.1f} %  waiting {100 * wait / total:5.1f} %  warp-instructions {ex / 1e6:7.2f} M")
    print(f"  total samples {int(total)}")


if __name__ == "__main__":
    main()
