#!/bin/bash
# ncu full capture of the Delta III kernel with the new scatter decode (why does the node phase stretch?)
bash tools/prof_generic.sh d3_new python tools/d3_eval.py 83333 4 | tail -1
tail -3 gpurun_out/ncu_d3_new.log
