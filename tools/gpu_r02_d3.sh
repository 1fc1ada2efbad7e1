#!/bin/bash
O=gpurun_out/r02_d3_knobs.txt; : > $O
python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_MIN_BLOCKS=3 python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_MIN_BLOCKS=2 python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_SMEM_BUDGET=55000 python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_SMEM_BUDGET=55000 PCX_MIN_BLOCKS=4 python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_SMEM_BUDGET=44000 PCX_MIN_BLOCKS=5 python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_THREADS=64 PCX_MIN_BLOCKS=7 python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_THREADS=64 PCX_MIN_BLOCKS=6 python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_THREADS=96 PCX_MIN_BLOCKS=4 python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_THREADS=96 PCX_MIN_BLOCKS=5 python tools/d3_eval.py 83333 10 >> $O 2>&1
grep '^{' $O | cut -c1-260
grep -v '^{' $O | tail -5
