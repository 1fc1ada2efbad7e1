#!/bin/bash
# Delta III two-pass: Hessian staged in output order (flush = straight copy); CTA sizes
O=gpurun_out/r02_d3_twopass2.txt; : > $O
PCX_NVRTC_EXTRA="-DPCX_TWO_PASS=1" python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_THREADS=96 PCX_NVRTC_EXTRA="-DPCX_TWO_PASS=1" python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_THREADS=64 PCX_NVRTC_EXTRA="-DPCX_TWO_PASS=1" python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_THREADS=96 python tools/d3_eval.py 83333 10 >> $O 2>&1
grep '^{' $O | cut -c1-300
grep -v '^{' $O | tail -5
PCX_NVRTC_EXTRA="-DPCX_TWO_PASS=1" python tools/d3_timeline.py 2>&1 | tail -12
PCX_NVRTC_EXTRA="-DPCX_TWO_PASS=1" timeout 900 python -m pytest tests/test_reference_goldens.py -m gpu -x -q -k "delta or robot or shuttle" 2>&1 | tail -3
