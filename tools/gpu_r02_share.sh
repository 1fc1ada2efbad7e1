#!/bin/bash
# one expression body shared by the phases that differ in literals only (codegen.share_groups): A/B + parity
O=gpurun_out/r02_d3_share.txt; : > $O
python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_SHARE_BODIES=0 python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_ALT_ORDER=1 python tools/d3_eval.py 83333 10 >> $O 2>&1
grep '^{' $O | cut -c1-300
grep -v '^{' $O | tail -5
timeout 400 python -m pytest tests -m gpu -q -k "delta or multiphase or sliding" 2>&1 | tail -3
python tools/d3_timeline.py 2>&1 | grep -A12 "first generation"
