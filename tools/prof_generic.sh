#!/bin/bash
# ncu full capture of the fused G+H kernel for an arbitrary tool command:
#   tools/prof_generic.sh TAG <command...>
TAG=$1; shift
O=gpurun_out
PCX_NO_GATE=1 PCX_DUMP_SRC=1 ncu --set full --clock-control none --import-source on -k regex:pcx_fill -s 3 -c 1 -f -o $O/prof_$TAG "$@" > $O/ncu_$TAG.log 2>&1
ncu -i $O/prof_$TAG.ncu-rep --page raw --csv > $O/raw_$TAG.csv 2>/dev/null
ncu -i $O/prof_$TAG.ncu-rep --page source --print-source cuda,sass --csv > $O/src_$TAG.csv 2>/dev/null
rm -f pcx_kernels.cu pcx_problem.h pcx_params.h
ls -la $O/prof_$TAG.ncu-rep
