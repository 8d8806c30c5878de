#!/bin/bash
# development aid: time the aggregation launches with and without the split-boundary sweep (agg_vsweep2_kernel)
for v in ${VARIANTS:-1 0}; do
  echo "== B2S_VSWEEP2=$v"
  B2S_VSWEEP2=$v timeout 300 python scripts/quick_timing.py 2>&1 | grep -E "aggregate alone|disp md5|aggregation launches|rror"
done
