#!/bin/bash
O=gpurun_out
# headline kernel: Hessian entries staged (full-sector stores) vs direct
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $O/bench_z_direct.json 2>$O/bench_z.err; python -c "
import json;d=json.loads(open('$O/bench_z_direct.json').read().strip().splitlines()[-1]);print('direct ',d['value'],d['ms_per_step'],d['ordered']['us_per_eval'])"
PCX_STAGE_H=1 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $O/bench_z_staged.json 2>$O/bench_z2.err; python -c "
import json;d=json.loads(open('$O/bench_z_staged.json').read().strip().splitlines()[-1]);print('staged ',d['value'],d['ms_per_step'],d['ordered']['us_per_eval'])"
tail -2 $O/bench_z2.err
# final Delta III kernel: full ncu capture + source page
bash tools/prof_generic.sh d3_final python tools/d3_eval.py 83333 4 | tail -1
