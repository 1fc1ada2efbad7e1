#!/bin/bash
# Delta III: first-wave start times spread uniformly over one tile lifetime (global de-phasing)
O=gpurun_out/r02_d3_stagger2.txt; : > $O
for ns in 10000 19000 30000; do
PCX_NVRTC_EXTRA="-DPCX_TWO_PASS=1 -DPCX_STAGGER_NS=$ns" python tools/d3_eval.py 83333 10 >> $O 2>&1
done
PCX_NVRTC_EXTRA="-DPCX_STAGGER_NS=22000" python tools/d3_eval.py 83333 10 >> $O 2>&1
grep '^{' $O | cut -c1-300
grep -v '^{' $O | tail -5
PCX_NVRTC_EXTRA="-DPCX_TWO_PASS=1 -DPCX_STAGGER_NS=19000" python tools/d3_timeline.py 2>&1 | tail -14
