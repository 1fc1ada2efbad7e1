#!/bin/bash
O=gpurun_out
for v in "" "-DPCX_PRE_STAGED=1"; do
PCX_NVRTC_EXTRA="$v" python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $O/bench_ac.json 2>$O/bench_ac.err; python -c "
import json;d=json.loads(open('$O/bench_ac.json').read().strip().splitlines()[-1]);print('[$v]',round(d['value']),d['ms_per_step'],'ordered',d['ordered']['us_per_eval'],'latency',d['latency_us']['value'])"
done
for v in "" "-DPCX_PRE_STAGED=1"; do
PCX_NVRTC_EXTRA="$v" python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_ac.json 2>$O/bench_ac.err; python -c "
import json;d=json.loads(open('$O/bench_ac.json').read().strip().splitlines()[-1]);print('[$v] 20 steps',round(d['value']),d['ms_per_step'],'ordered',d['ordered']['us_per_eval'],'latency',d['latency_us']['value'])"
done
