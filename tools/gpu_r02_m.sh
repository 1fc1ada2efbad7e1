#!/bin/bash
# Delta III after the residency fix (register cap follows the shared-memory-limited residency)
O=gpurun_out/r02_d3_decode_ab2.txt; : > $O
python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_NVRTC_EXTRA="-DPCX_DECODE_V1=1" python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_NVRTC_EXTRA="-DPCX_EARLY_WAIT=1" python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_NVRTC_EXTRA="-DPCX_SCATTER_UNROLL=8" python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_STAGE_H=1 python tools/d3_eval.py 83333 10 >> $O 2>&1
grep '^{' $O | cut -c1-300
grep -v '^{' $O | tail -5
python tools/d3_timeline.py > gpurun_out/r02_d3_timeline_new2.txt 2>&1; tail -12 gpurun_out/r02_d3_timeline_new2.txt
