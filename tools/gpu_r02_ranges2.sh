#!/bin/bash
# one GPU standing in for each rank of an 8-way sharded Delta III mesh, kernel with the shared expression body
O=gpurun_out/r02_ranges_share.txt; : > $O
timeout 125 python tools/range_time.py 8 83333 >> $O 2>&1
grep '^{' $O | cut -c1-400
grep -v '^{' $O | tail -3
