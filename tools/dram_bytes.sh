#!/bin/bash
# Dev tool: DRAM bytes of the backward kernel (3rd k_gemm_tc launch of a step) for given debug builds (run under gpurun)
for l in "$@"; do
  echo "== $l"
  TGB_DBG_LIB=$l ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:${KERN:-k_gemm_tc_pair} --launch-skip ${SKIP:-1} --launch-count 1 python tools/mainloop_only.py 2>&1 | grep -E "dram__|gpu__time|==ERROR|rror" | cut -c1-200
done
