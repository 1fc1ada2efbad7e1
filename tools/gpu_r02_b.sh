#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --deselect tests/test_sharding.py::test_gpu_fused_exchange_across_processes -s > gpurun_out/r02b_gputests.log 2>&1
echo "pytest exit $?"; tail -25 gpurun_out/r02b_gputests.log | cut -c1-300
python tools/knobs.py default_t192 > gpurun_out/r02b_knobs.txt 2>&1; grep '^{' gpurun_out/r02b_knobs.txt | cut -c1-250
