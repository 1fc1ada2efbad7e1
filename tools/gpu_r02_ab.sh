#!/bin/bash
O=gpurun_out/r02_d3_prologue.txt; : > $O
python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_NVRTC_EXTRA="-DPCX_PRE_STAGED=1" python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_NVRTC_EXTRA="-DPCX_PRE_STAGED=1 -DPCX_EARLY_WAIT=1" python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_TILES_PER_SM=48 python tools/d3_eval.py 83333 10 >> $O 2>&1
grep '^{' $O | cut -c1-300
grep -v '^{' $O | tail -5
