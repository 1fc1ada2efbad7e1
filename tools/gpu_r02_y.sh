#!/bin/bash
# validation pass after the Delta III work: full GPU suite, headline bench at the driver's settings, large-body timings
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/r02_y_tests.txt
python bench.py --steps 20 --warmup 5 > $O/bench_r02_y.json 2> $O/bench_r02_y.err; cut -c1-700 $O/bench_r02_y.json; tail -2 $O/bench_r02_y.err
python tools/d3_eval.py 83333 10 > $O/r02_y_d3.txt 2>&1; tail -1 $O/r02_y_d3.txt
python tools/d3_eval.py 333333 10 space_shuttle_reentry >> $O/r02_y_d3.txt 2>&1; tail -1 $O/r02_y_d3.txt
python tools/d3_eval.py 200000 10 free_flying_robot >> $O/r02_y_d3.txt 2>&1; tail -1 $O/r02_y_d3.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
