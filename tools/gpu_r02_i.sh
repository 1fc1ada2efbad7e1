#!/bin/bash
# Delta III scatter: staged run tables + prefetched recipe words; early dependency wait (A/B)
O=gpurun_out/r02_d3_decode.txt; : > $O
python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_NVRTC_EXTRA="-DPCX_EARLY_WAIT=1" python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_NVRTC_EXTRA="-DPCX_EARLY_WAIT=1 -DPCX_SCATTER_UNROLL=8" python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_NVRTC_EXTRA="-DPCX_SCATTER_UNROLL=2" python tools/d3_eval.py 83333 10 >> $O 2>&1
grep '^{' $O | cut -c1-300
grep -v '^{' $O | tail -5
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02_i_tests.txt
