#!/bin/bash
O=gpurun_out/r02_d3_altorder.txt; : > $O
python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_ALT_ORDER=1 python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_ALT_ORDER=1 python tools/d3_eval.py 333333 10 space_shuttle_reentry >> $O 2>&1
grep '^{' $O | cut -c1-300
grep -v '^{' $O | tail -5
PCX_ALT_ORDER=1 python tools/d3_timeline.py 2>&1 | grep -A12 "first generation"
PCX_ALT_ORDER=1 timeout 300 python -m pytest tests/test_reference_goldens.py -m gpu -x -q -k "delta" 2>&1 | tail -2
