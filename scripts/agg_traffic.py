"""Regenerate profiles/agg_traffic.json from an `ncu --set full` capture of the aggregation group (same commit as the library):

    # on the GPU box (gpurun), one pair's aggregation launches of the production schedule:
    ncu --set full --clock-control none --import-source on -k regex:'agg_|wta_kernel' -s <first launch of the third call> -c <n> \
        -o gpurun_out/r02_agg python scripts/quick_timing.py > gpurun_out/r02_agg.log
    # here:
    python scripts/agg_traffic.py gpurun_out/r02_agg.ncu-rep gpurun_out/r02_agg.log [profiles/r02_agg_raw.csv]

The log must contain the line `build_hash <16 hex digits>` (scripts/quick_timing.py prints b2s_build_hash()); bench.py uses
the traffic figure only when that hash equals the hash of the library it runs."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "usecond": 1e-3, "msecond": 1.0, "second": 1e3, "nsecond": 1e-6, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}


def main():
    rep, log = sys.argv[1], sys.argv[2]
    raw_out = sys.argv[3] if len(sys.argv) > 3 else None
    m = re.search(r"build_hash ([0-9a-f]{16})", open(log).read())
    if not m:
        raise SystemExit("no `build_hash` line in %s" % log)
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    if raw_out:
        open(raw_out, "w").write(txt)
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {k: i for i, k in enumerate(hdr)}

    def val(r, k):
        return float(r[col[k]].replace(",", "")) * UNIT.get(units[col[k]], 1)

    kernels, seen = [], set()
    for r in body:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void <unnamed>::", "")
        if name in seen:  # the first pair's launches only
            break
        seen.add(name)
        kernels.append({"kernel": name, "ms": val(r, "gpu__time_duration.sum"), "dram_bytes_read": val(r, "dram__bytes_read.sum"),
                        "dram_bytes_write": val(r, "dram__bytes_write.sum"),
                        "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                        "alu_pipe_pct": val(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                        "inst_executed": val(r, "smsp__inst_executed.sum"), "registers": val(r, "launch__registers_per_thread")})
    total = sum(k["dram_bytes_read"] + k["dram_bytes_write"] for k in kernels)
    out = {"build_hash": m.group(1), "source": os.path.basename(rep), "launches_per_pair": len(kernels), "dram_bytes_per_pair": total,
           "dram_bytes_per_launch": total / max(len(kernels), 1), "kernels": kernels,
           "note": "ncu --set full --clock-control none, one pair at 1080p / 128 disparities MODE_HH; per-launch times under ncu are cold-cache and serialised"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "agg_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
