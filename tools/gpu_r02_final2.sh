#!/bin/bash
# last validation pass of round 2 (one GPU): all GPU tests, the bench at the driver's settings, smoke,
# and -- if the remaining budget allows -- one full ncu capture of the Delta III kernel with the shared body
O=gpurun_out; mkdir -p $O
SECONDS=0
timeout 280 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/r02_final2_tests.txt
echo "tests done at ${SECONDS}s"
timeout 150 python bench.py --steps 20 --warmup 5 > $O/bench_r02_final2.json 2> $O/bench_r02_final2.err; cut -c1-200 $O/bench_r02_final2.json; tail -2 $O/bench_r02_final2.err
echo "bench done at ${SECONDS}s"
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
if [ $SECONDS -lt 300 ]; then
  timeout $((385 - SECONDS)) bash tools/prof_generic.sh d3_share python tools/d3_eval.py 83333 4 | tail -1
fi
echo "all done at ${SECONDS}s"
