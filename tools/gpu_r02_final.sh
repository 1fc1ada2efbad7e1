#!/bin/bash
# final validation pass of round 2 (one GPU): tests, benches, captures for profiles/
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/r02_final_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > $O/bench_r02_final.json 2> $O/bench_r02_final.err; cut -c1-200 $O/bench_r02_final.json; tail -2 $O/bench_r02_final.err
python bench.py > $O/bench_r02_final_default.json 2> $O/bench_r02_final_default.err; cut -c1-200 $O/bench_r02_final_default.json
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref_r02_final.json 2> $O/bench_ref_r02_final.err; cut -c1-300 $O/bench_ref_r02_final.json
python tools/d3_eval.py 83333 10 > $O/r02_final_d3.txt 2>&1
python tools/d3_eval.py 333333 10 space_shuttle_reentry >> $O/r02_final_d3.txt 2>&1
python tools/d3_eval.py 200000 10 free_flying_robot >> $O/r02_final_d3.txt 2>&1
grep '^{' $O/r02_final_d3.txt | cut -c1-200
python tools/d3_timeline.py > $O/r02_final_d3_timeline.txt 2>&1; sed -n '1,14p' $O/r02_final_d3_timeline.txt
bash tools/prof_generic.sh d3_final python tools/d3_eval.py 83333 4 | tail -1
python tools/adapter_bench.py > $O/r02_adapter3.txt 2>&1; tail -3 $O/r02_adapter3.txt | cut -c1-250
