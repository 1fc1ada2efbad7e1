#!/bin/bash
# Delta III: de-phasing the CTAs that share an SM (first-wave stagger)
O=gpurun_out/r02_d3_stagger.txt; : > $O
for ns in 0 3000 6000 9000; do
PCX_NVRTC_EXTRA="-DPCX_TWO_PASS=1 -DPCX_STAGGER_NS=$ns" python tools/d3_eval.py 83333 10 >> $O 2>&1
done
PCX_NVRTC_EXTRA="-DPCX_STAGGER_NS=7000" python tools/d3_eval.py 83333 10 >> $O 2>&1
grep '^{' $O | cut -c1-300
grep -v '^{' $O | tail -5
PCX_NVRTC_EXTRA="-DPCX_TWO_PASS=1" python tools/d3_timeline.py 2>&1 | tail -14
PCX_NVRTC_EXTRA="-DPCX_TWO_PASS=1 -DPCX_STAGGER_NS=6000" python tools/d3_timeline.py 2>&1 | tail -14
