#!/bin/bash
# Hessian-only launches of large bodies: staged full-sector flush vs direct stores; correctness of the separate variants
O=gpurun_out/r02_hst.txt; : > $O
WHAT=h python tools/d3_eval.py 83333 10 >> $O 2>&1
WHAT=h PCX_NVRTC_EXTRA="-DPCX_TWO_PASS=0" python tools/d3_eval.py 83333 10 >> $O 2>&1
WHAT=j python tools/d3_eval.py 83333 10 >> $O 2>&1
grep '^{' $O | cut -c1-300
grep -v '^{' $O | tail -5
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
