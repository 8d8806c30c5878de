"""SASS evidence for profiles/: per hot kernel of libb2s.so the mnemonic histogram of `cuobjdump -sass` and the count of the
Blackwell-relevant instructions (DPX VIMNMX3 / VIADDMNMX, CREDUX, bulk copies UBLKCP, mbarrier SYNCS, cp.async LDGSTS, IDP).
    python scripts/sass_summary.py > profiles/r02_sass_summary.txt          (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from calibrating_b200 import build  # noqa: E402

HOT = ["agg_vsweep2_kernelILi2ELb0ELi8ELb0E", "agg_vsweep_kernelILi2ELb0ELi2ELi8E", "agg_hscan_kernelILi2ELb0ELi2ELi1ELb0E", "agg_hscan_vsum_kernelILi2ELb0ELi5ELb0E", "pixcost_hsum_kernelILi3ELi104ELi0E",
       "agg_wave_kernelILi2ELb0E", "agg_vsweep6_kernelILb0E", "remap_lz4_kernelILi3ELb1E", "wta_kernelILi2ELb1ELi1E", "planes_kernelILi3E", "lr_median_kernel",
       "resize_u8_kernelILi3E", "cloud_emit_kernelILi3E"]
KEY = ["VIMNMX3", "VIADDMNMX", "VIMNMX", "CREDUX", "UBLKCP", "SYNCS", "LDGSTS", "IDP", "SHFL", "PRMT", "LDS", "STS", "LDG", "STG", "ATOM", "RED", "BAR", "DADD", "DMUL", "DFMA"]


def main():
    txt = subprocess.check_output(["cuobjdump", "-sass", build.OUT], text=True)
    funcs, name = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            ins = re.sub(r"/\*.*?\*/", "", line).strip().rstrip(";").strip()
            ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
            if ins:
                funcs[name].append(ins.split()[0])
    print("libb2s.so build %s (sha256 over the CUDA sources, first 16 hex digits); cuobjdump -sass, sm_100a" % build.source_hash())
    total = collections.Counter()
    for ops in funcs.values():
        total.update(o.split(".")[0] for o in ops)
    print("whole library: %d kernels, %d instructions; %s" % (len(funcs), sum(total.values()), ", ".join("%s %d" % (k, total[k]) for k in KEY if total[k])))
    print()
    for h in HOT:
        for fn, ops in funcs.items():
            if h in fn:
                c = collections.Counter(o.split(".")[0] for o in ops)
                full = collections.Counter(ops)
                print("%s\n  %d instructions; %s" % (fn, len(ops), ", ".join("%s %d" % (k, c[k]) for k in KEY if c[k])))
                print("  most frequent: " + ", ".join("%s %d" % kv for kv in full.most_common(14)))
                break
        else:
            print("%s: not found" % h)


if __name__ == "__main__":
    main()
