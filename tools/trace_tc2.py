"""Timeline of one CTA of conv_tc2_kernel (developer tool): which role waits for which, per tile.

    python tools/trace_tc2.py N H W Cin Cout [K]

Installs the debug buffer (mn_debug_tc2_trace), runs the conv once, prints per-tile event times (in cycles relative to the first
event) for CTA 0:  1 halo TMA issue, 2 halo landed, 3 split done, 4 feed starts, 5 feed done, 6 accumulator free, 7 last MMA
issued, 8 accumulators complete, 9 drain done, 10 tile stored."""
import ctypes
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marconet_b200 import _lib, ops  # noqa: E402

NAMES = {1: "halo_issue", 2: "halo_landed", 3: "split_done", 4: "feed_start", 5: "feed_done", 6: "acc_free", 7: "mma_issued",
         8: "acc_complete", 9: "drain_done", 10: "stored", 11: "staged", 12: "rows_done", 13: "epi_top", 14: "a_free", 15: "st_done"}
MAIN = list(range(1, 11))


def main():
    n, h, w, cin, cout = [int(v) for v in sys.argv[1:6]]
    k = int(sys.argv[6]) if len(sys.argv) > 6 else 3
    dev = torch.device("cuda:0")
    lib = _lib.load()
    fn = lib.mn_debug_tc2_trace
    fn.restype, fn.argtypes = ctypes.c_int, [ctypes.c_void_p]
    x = torch.randn(n, h, w, cin, device=dev)
    wt = ops.ConvWeight((torch.randn(k * k * cin, cout, device=dev) / (k * k * cin) ** 0.5).contiguous(), k * k)
    bias = torch.randn(cout, device=dev)
    for _ in range(3):
        ops.conv2d(x, wt, k, k, pad=(k // 2, k // 2), bias=bias, act=ops.ACT_LRELU02, gain=2 ** 0.5)
    buf = torch.zeros(1 + 64 * 64, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    fn(ctypes.c_void_p(buf.data_ptr()))
    ops.conv2d(x, wt, k, k, pad=(k // 2, k // 2), bias=bias, act=ops.ACT_LRELU02, gain=2 ** 0.5)
    torch.cuda.synchronize()
    fn(None)
    host = buf.cpu().tolist()
    t0 = min(v for v in host[1:] if v)
    per = defaultdict(lambda: defaultdict(list))
    cnt = 0
    for it in range(64):
        for e in range(1, 64):
            v = host[1 + it * 64 + e]
            if v:
                cnt += 1
                per[it][e if e < 16 else (14 if e < 40 else 15)].append(v - t0)
    print(f"conv N{n} {h}x{w} {cin}->{cout} k{k}: {cnt} events, CTA 0, cycles since first event; rows = tile iterations of this CTA")
    works = sorted(per)
    print("work  " + "  ".join(f"{NAMES[e]:>13s}" for e in MAIN))
    prev_stored = None
    for wk in works:
        row = []
        for e in MAIN:
            ts = per[wk].get(e, [])
            row.append(f"{ts[0]:>6d}..{ts[-1]:<6d}" if len(ts) > 1 else (f"{ts[0]:>13d}" if ts else " " * 13))
        print(f"{wk:4d}  " + "  ".join(row))
    stored = [per[wk][10][0] for wk in works if per[wk].get(10)]
    if len(stored) > 2:
        gaps = [b - a for a, b in zip(stored, stored[1:])]
        print("tile period (stored -> stored), cycles:", gaps)
    mid = works[len(works) // 2]
    print(f"detail of work {mid}: epilogue top/acc_complete/drain_done/staged/rows_done/stored:",
          [per[mid].get(e, [None])[0] for e in (13, 8, 9, 11, 12, 10)])
    print("  feed per tap: A stage free", per[mid].get(14), "\n  TMEM stores complete", per[mid].get(15))


if __name__ == "__main__":
    main()
