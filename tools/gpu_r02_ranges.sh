#!/bin/bash
# one GPU standing in for each rank of an 8-way / 4-way sharded Delta III mesh: final kernel vs the first-half kernel
O=gpurun_out/r02_ranges_final.txt; : > $O
python tools/range_time.py 8 83333 >> $O 2>&1
PCX_NVRTC_EXTRA="-DPCX_STORE_TOKEN=0" python tools/range_time.py 8 83333 >> $O 2>&1
PCX_NVRTC_EXTRA="-DPCX_TWO_PASS=0 -DPCX_STORE_TOKEN=0" python tools/range_time.py 8 83333 >> $O 2>&1
python tools/range_time.py 4 83333 >> $O 2>&1
grep '^{' $O | cut -c1-400
grep -v '^{' $O | tail -3
