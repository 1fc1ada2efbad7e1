#!/bin/bash
O=gpurun_out/r02_stageh.txt; : > $O
PCX_STAGE_H=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_sharding.py -m gpu -q -x --deselect tests/test_sharding.py::test_gpu_fused_exchange_across_processes 2>&1 | tail -4
PCX_STAGE_H=0 python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_STAGE_H=1 python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_STAGE_H=1 PCX_SMEM_BUDGET=72000 python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_STAGE_H=1 PCX_SMEM_BUDGET=56000 python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_STAGE_H=1 PCX_SMEM_BUDGET=72000 PCX_THREADS=96 python tools/d3_eval.py 83333 10 >> $O 2>&1
PCX_STAGE_H=1 PCX_SMEM_BUDGET=50000 PCX_THREADS=64 python tools/d3_eval.py 83333 10 >> $O 2>&1
grep '^{' $O | cut -c1-300; grep -v '^{' $O | tail -3
