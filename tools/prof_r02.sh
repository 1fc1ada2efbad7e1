#!/bin/bash
# round-2 profiling pass (one GPU): tests, bench, ncu captures for profiles/
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -s --deselect tests/test_sharding.py::test_gpu_fused_exchange_across_processes > $O/r02_gputests.log 2>&1
echo "pytest exit $?"; tail -4 $O/r02_gputests.log | cut -c1-300
grep -n "^solution chain\|^cuda " $O/r02_gputests.log | cut -c1-200 | tail -22
# 1) cart-pole headline kernel: bench, full capture, launch list   (tools/prof.sh)
bash tools/prof.sh r02 | tail -3 | cut -c1-400
# 2) steady-state DRAM bytes: 12 consecutive ring launches in their natural cache state
PCX_NO_GATE=1 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    -k regex:pcx_fill -s 12 -c 12 --csv --log-file $O/steady_r02.csv python bench.py --steps 40 --warmup 5 --no-cpu-baseline > $O/ncu_steady_r02.log 2>&1
tail -3 $O/steady_r02.csv | cut -c1-300
# 3) Delta III, 10^6 nodes, one GPU: timing + full capture
python tools/d3_eval.py > $O/d3_r02.txt 2>&1; tail -1 $O/d3_r02.txt
bash tools/prof_generic.sh d3_r02 python tools/d3_eval.py 83333 4 | tail -1
python tools/adapter_bench.py > $O/r02_adapter.txt 2>&1; tail -4 $O/r02_adapter.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > $O/bench_r02_driver.json 2> $O/bench_r02_driver.err; cut -c1-300 $O/bench_r02_driver.json
python bench.py > $O/bench_r02_default.json 2> $O/bench_r02_default.err; cut -c1-200 $O/bench_r02_default.json
python tools/config1.py > $O/r02_config1.txt 2>&1; tail -3 $O/r02_config1.txt
python tools/mesh_error_bench.py > $O/r02_mesh_error.txt 2>&1; tail -4 $O/r02_mesh_error.txt
python tools/multistart.py > $O/r02_multistart.txt 2>&1; tail -3 $O/r02_multistart.txt
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref_r02.json 2> $O/bench_ref_r02.err; cut -c1-400 $O/bench_ref_r02.json
