#!/bin/bash
# sharded Delta III ranges on one GPU: fewest tiles the node cap allows against whole waves
O=gpurun_out/r02_d3_tiling2.txt; : > $O
timeout 52 python tools/tiling_ranges.py 8:7 8:8 4:14 4:16 2:27 2:28 >> $O 2>&1
grep '^{' $O | cut -c1-300
grep -v '^{' $O | tail -3
