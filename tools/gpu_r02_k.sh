#!/bin/bash
# Delta III: new decode with the Hessian entries staged (coalesced, full sectors) -- store-sector pressure on the SM->L2 path
O=gpurun_out/r02_d3_stageh.txt; : > $O
PCX_STAGE_H=1 python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_STAGE_H=1 PCX_SMEM_BUDGET=56000 python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_STAGE_H=1 PCX_SMEM_BUDGET=72000 python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_STAGE_H=1 PCX_SMEM_BUDGET=110000 python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_STAGE_H=1 PCX_NVRTC_EXTRA="-DPCX_DECODE_V1=1" python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_SMEM_BUDGET=110000 python tools/d3_eval.py 83333 10 >> $O 2>&1
grep '^{' $O | cut -c1-300
grep -v '^{' $O | tail -5
PCX_STAGE_H=1 python tools/d3_timeline.py > gpurun_out/r02_d3_timeline_stageh.txt 2>&1; tail -12 gpurun_out/r02_d3_timeline_stageh.txt
