#!/bin/bash
# Run on the GPU box (via gpurun): bench, then ncu launch list + one full capture
# of the fused G+H kernel with per-line source counters exported as CSV.
#   tools/prof.sh TAG
TAG=${1:-x}
O=gpurun_out
python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_$TAG.json 2> $O/bench_$TAG.err || { tail -5 $O/bench_$TAG.err; exit 1; }
cat $O/bench_$TAG.json
PCX_NO_GATE=1 PCX_DUMP_SRC=1 ncu --set full --clock-control none --import-source on -k regex:pcx_fill -s 8 -c 1 -f -o $O/prof_$TAG \
    python bench.py --steps 12 --warmup 4 --no-cpu-baseline > $O/ncu_$TAG.log 2>&1
ncu -i $O/prof_$TAG.ncu-rep --page raw --csv > $O/raw_$TAG.csv 2>/dev/null
ncu -i $O/prof_$TAG.ncu-rep --page source --print-source cuda,sass --csv > $O/src_$TAG.csv 2>/dev/null
ls -la $O/prof_$TAG.ncu-rep
rm -f pcx_kernels.cu pcx_problem.h pcx_params.h
# launch list of the same command (per-launch durations, cold-cache and serialised)
PCX_NO_GATE=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/ncu_launch_$TAG.log 2>&1
