#!/bin/bash
# Delta III: new decode vs round-1 decode on the SAME box, per-CTA timelines; adapter with registered uploads
O=gpurun_out/r02_d3_decode_ab.txt; : > $O
python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_NVRTC_EXTRA="-DPCX_DECODE_V1=1" python tools/d3_eval.py 83333 10 >> $O 2>&1
python tools/d3_eval.py 83333 10 >> $O 2>&1
grep '^{' $O | cut -c1-300
grep -v '^{' $O | tail -5
python tools/d3_timeline.py > gpurun_out/r02_d3_timeline_new.txt 2>&1; tail -12 gpurun_out/r02_d3_timeline_new.txt
PCX_NVRTC_EXTRA="-DPCX_DECODE_V1=1" python tools/d3_timeline.py > gpurun_out/r02_d3_timeline_v1.txt 2>&1; tail -12 gpurun_out/r02_d3_timeline_v1.txt
timeout 600 python -m pytest tests/test_gpu_edges.py -x -q -k adapter 2>&1 | tail -15
python tools/adapter_bench.py > gpurun_out/r02_adapter2.txt 2>&1; cat gpurun_out/r02_adapter2.txt | tail -8
