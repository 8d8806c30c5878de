"""Row timeline of agg_vsweep2_kernel (development aid).  Run under gpurun:

    B2S_VS2_TRACE=gpurun_out/vs2_trace.bin python scripts/quick_timing.py
    python scripts/vs2_trace.py gpurun_out/vs2_trace.bin

The TRACE instantiation stores clock64 at fixed points of rows 512..527 for every warp of CTAs 60..62 (same SM clock inside a CTA).
Points -- interior column: 0 loop top, 1 neighbours' states available, 2 own states handed over, 3 sums stored, 4 next row loaded;
boundary column: 0 top, 1 inner neighbour's state available, 2 outgoing diagonal written to the neighbour CTA, 3 incoming diagonal
received, 4 handed to the inner neighbour, 5 sums passed to the helper, 6 next row loaded; helper: 0 top, 1 vertical step done, 2 the
column warp's sums available, 3 stored, 4 next row loaded."""
import sys
import numpy as np

t = np.fromfile(sys.argv[1], dtype=np.int64).reshape(3, 32, 16, 8)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 13
names = {}
for wi in range(2 * n + 4):
    if wi < 8:
        jl, w, helper = (wi >> 1) & 1, (n - 1 if wi & 1 else 0), wi >= 4
        names[wi] = "%s %s col %2d" % ("down" if jl == 0 else "up  ", "helper" if helper else "edge  ", w)
    else:
        jl, w = (wi - 8) // (n - 2), 1 + (wi - 8) % (n - 2)
        names[wi] = "%s        col %2d" % ("down" if jl == 0 else "up  ", w)
for cta in range(3):
    base = t[cta][t[cta] > 0].min() if (t[cta] > 0).any() else 0
    print("== CTA %d (cycles relative to the first mark of the CTA)" % (60 + cta))
    order = sorted(names, key=lambda wi: (names[wi][:4], int(names[wi][-2:]), "helper" in names[wi]))
    for wi in order:
        rows = t[cta, wi]
        if not (rows > 0).any():
            continue
        per = np.diff(rows[:, 0][rows[:, 0] > 0])
        seg = []
        for r in range(4, 8):
            m = rows[r]
            pts = [int(v - base) if v > 0 else None for v in m]
            seg.append(" ".join("%6s" % ("-" if v is None else v) for v in pts[:7]))
        dur = []
        npts = 7 if "edge" in names[wi] else 5
        for k in range(npts - 1):
            ok = (rows[:, k] > 0) & (rows[:, k + 1] > 0)
            dur.append(int(np.mean(rows[ok, k + 1] - rows[ok, k])) if ok.any() else -1)
        print("%s | period %5d | mean segment cycles %s" % (names[wi], int(per.mean()) if len(per) else -1, dur))
    print()
