#!/bin/bash
O=gpurun_out/r02_knobs_indep.txt; : > $O
python tools/knobs.py t192 >> $O 2>&1
PCX_THREADS=128 python tools/knobs.py t128 >> $O 2>&1
PCX_THREADS=256 python tools/knobs.py t256 >> $O 2>&1
PCX_THREADS=160 python tools/knobs.py t160 >> $O 2>&1
PCX_TILES_PER_SM=6 python tools/knobs.py t192_tiles6 >> $O 2>&1
PCX_TILES_PER_SM=8 python tools/knobs.py t192_tiles8 >> $O 2>&1
PCX_THREADS=128 PCX_TILES_PER_SM=9 python tools/knobs.py t128_tiles9 >> $O 2>&1
PCX_THREADS=128 PCX_TILES_PER_SM=12 python tools/knobs.py t128_tiles12 >> $O 2>&1
grep '^{' $O | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['label'], d['tiles'], d['threads'], 'ordered', d['us_per_eval_200'], d['us_per_eval_20'], 'indep', d['us_independent_200'], d['us_independent_20'], d['independent_equal'])"
