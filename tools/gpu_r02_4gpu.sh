#!/bin/bash
# 2 GPUs: the cross-process fused-exchange test and bench.py --gpus 4 (strong object) with the final kernel
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_bench_4gpu.json 2> gpurun_out/r02_bench_4gpu.err
echo "bench exit $?"; grep -v "^\*\|OMP" gpurun_out/r02_bench_4gpu.err | tail -5 | cut -c1-400
python - <<'PY'
import json
s=open("gpurun_out/r02_bench_4gpu.json").read()
d=json.loads(s[s.index('{"metric"'):])
print("value", d["value"], "frac/gpu", d["roofline"]["frac"]/4, "e2e", d["e2e"]["value"])
print(json.dumps(d.get("strong"), indent=1)[:1800])
PY
