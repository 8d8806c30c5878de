#!/bin/bash
# development aid: what-if timings of agg_vsweep2_kernel at 1080p/128.  B2S_VS2_FAKE bits: 1 = no polling of the
# hand-over rings, 2 = no neighbour waits, 4 = hand-over prefetch issued at the start of the next row; results are WRONG when 1 or 2 is set.
# B2S_VSWEEP2=0 = the round-1 kernel (agg_vsweep_kernel).
for v in ${VARIANTS:-0 4 1 2 3}; do
  echo "== B2S_VS2_FAKE=$v"
  B2S_VS2_FAKE=$v timeout 300 python scripts/quick_timing.py 2>&1 | grep -E "disp md5|aggregation launches|rror"
done
echo "== B2S_VSWEEP2=0"
B2S_VSWEEP2=0 timeout 300 python scripts/quick_timing.py 2>&1 | grep -E "disp md5|aggregation launches|rror"
