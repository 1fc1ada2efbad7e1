#!/bin/bash
# Delta III knock-outs (timing only, wrong results): what do the Hessian stores / the Jacobian stores cost in steady state?
O=gpurun_out/r02_d3_knockout.txt; : > $O
python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_NVRTC_EXTRA="-DPCX_KO_HST=1" python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_NVRTC_EXTRA="-DPCX_KO_GST=1" python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_NVRTC_EXTRA="-DPCX_KO_HST=1 -DPCX_KO_GST=1" python tools/d3_eval.py 83333 10 >> $O 2>&1
grep '^{' $O | cut -c1-300
grep -v '^{' $O | tail -5
PCX_NVRTC_EXTRA="-DPCX_KO_HST=1" python tools/d3_timeline.py 2>&1 | tail -9
PCX_NVRTC_EXTRA="-DPCX_KO_GST=1" python tools/d3_timeline.py 2>&1 | tail -9
