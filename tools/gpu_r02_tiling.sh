#!/bin/bash
# Delta III tiling with the corrected residency estimate against the previous one, and the 8-way ranges
O=gpurun_out/r02_d3_tiling.txt; : > $O
timeout 100 python tools/tiling_ab.py >> $O 2>&1
grep '^{' $O | cut -c1-400
grep -v '^{' $O | tail -4
