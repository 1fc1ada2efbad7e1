#!/bin/bash
# Delta III: two-pass node phase (Hessian entries staged over the vacated first-derivative staging, full-sector stores)
O=gpurun_out/r02_d3_twopass.txt; : > $O
python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_NVRTC_EXTRA="-DPCX_TWO_PASS=1" python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_NVRTC_EXTRA="-DPCX_TWO_PASS=1 -DPCX_PRE_STAGED=1" python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_NVRTC_EXTRA="-DPCX_PRE_STAGED=1" python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_NVRTC_EXTRA="-DPCX_TWO_PASS=1 -DPCX_PRE_STAGED=1 -DPCX_EARLY_WAIT=1" python tools/d3_eval.py 83333 10 >> $O 2>&1
grep '^{' $O | cut -c1-300
grep -v '^{' $O | tail -5
PCX_NVRTC_EXTRA="-DPCX_TWO_PASS=1" python tools/d3_timeline.py 2>&1 | tail -9
# correctness of the two-pass form: Delta III + ragged cases through the C ABI against the goldens / oracle
PCX_NVRTC_EXTRA="-DPCX_TWO_PASS=1 -DPCX_PRE_STAGED=1" timeout 900 python -m pytest tests/test_reference_goldens.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
