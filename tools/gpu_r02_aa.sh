#!/bin/bash
O=gpurun_out/r02_d3_copyloop.txt; : > $O
python tools/d3_eval.py 83333 10 >> $O 2>&1
python tools/d3_eval.py 333333 10 space_shuttle_reentry >> $O 2>&1
PCX_NVRTC_EXTRA="-DPCX_STORE_TOKEN=0" python tools/d3_eval.py 83333 10 >> $O 2>&1
grep '^{' $O | cut -c1-300
grep -v '^{' $O | tail -5
python tools/d3_timeline.py 2>&1 | sed -n '1,14p'
timeout 600 python -m pytest tests/test_reference_goldens.py -m gpu -x -q -k "delta or shuttle" 2>&1 | tail -3
