#!/bin/bash
# round 2, first GPU contact: tests, bench at the driver's settings and defaults, knobs
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --deselect tests/test_sharding.py::test_gpu_fused_exchange_across_processes > gpurun_out/r02a_gputests.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/r02a_gputests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02a_bench_20.json 2> gpurun_out/r02a_bench_20.err
echo "bench20 exit $?"; cat gpurun_out/r02a_bench_20.json | cut -c1-600
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02a_bench_20b.json 2>> gpurun_out/r02a_bench_20.err
python bench.py --steps 1000 --warmup 20 --no-cpu-baseline > gpurun_out/r02a_bench_1000.json 2>> gpurun_out/r02a_bench_20.err
for f in gpurun_out/r02a_bench_20b.json gpurun_out/r02a_bench_1000.json; do python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], d["value"], d["roofline"]["frac"], d["latency_us"], d["e2e"]["value"])
PY
done
python tools/knobs.py default > gpurun_out/r02a_knobs.txt 2>&1
PCX_TILES_PER_SM=9 python tools/knobs.py tiles9 >> gpurun_out/r02a_knobs.txt 2>&1
PCX_TILES_PER_SM=12 python tools/knobs.py tiles12 >> gpurun_out/r02a_knobs.txt 2>&1
PCX_TILES_PER_SM=12 PCX_THREADS=64 python tools/knobs.py tiles12_t64 >> gpurun_out/r02a_knobs.txt 2>&1
PCX_THREADS=192 python tools/knobs.py t192 >> gpurun_out/r02a_knobs.txt 2>&1
PCX_NVRTC_EXTRA="-DPCX_SCATTER_UNROLL=8" python tools/knobs.py unroll8 >> gpurun_out/r02a_knobs.txt 2>&1
PCX_NVRTC_EXTRA="-DPCX_SCATTER_UNROLL=2" python tools/knobs.py unroll2 >> gpurun_out/r02a_knobs.txt 2>&1
PCX_MIN_BLOCKS=8 python tools/knobs.py mb8 >> gpurun_out/r02a_knobs.txt 2>&1
PCX_MIN_BLOCKS=4 python tools/knobs.py mb4 >> gpurun_out/r02a_knobs.txt 2>&1
PCX_NO_PDL=1 python tools/knobs.py nopdl >> gpurun_out/r02a_knobs.txt 2>&1
grep '^{' gpurun_out/r02a_knobs.txt | cut -c1-260
python tools/adapter_bench.py > gpurun_out/r02a_adapter.txt 2>&1; cat gpurun_out/r02a_adapter.txt | tail -6
