#!/bin/bash
O=gpurun_out/r02_d3_token.txt; : > $O
python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_NVRTC_EXTRA="-DPCX_STORE_TOKEN=1" python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_NVRTC_EXTRA="-DPCX_STORE_TOKEN=1 -DPCX_TWO_PASS=0" python tools/d3_eval.py 83333 10 >> $O 2>&1
grep '^{' $O | cut -c1-300
grep -v '^{' $O | tail -5
PCX_NVRTC_EXTRA="-DPCX_STORE_TOKEN=1" python tools/d3_timeline.py 2>&1 | tail -16
