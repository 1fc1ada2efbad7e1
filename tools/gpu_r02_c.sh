#!/bin/bash
# 2 GPUs: cross-process fused exchange test + bench with the strong-scaling leg
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_sharding.py -m gpu -q -s -k "across_processes or fused" > gpurun_out/r02c_sharding.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/r02c_sharding.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02c_bench_2gpu.json 2> gpurun_out/r02c_bench_2gpu.err
echo "bench exit $?"; tail -3 gpurun_out/r02c_bench_2gpu.err | cut -c1-600
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02c_bench_2gpu.json"))
print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
print(json.dumps(d.get("strong"), indent=1)[:1800])
PY
