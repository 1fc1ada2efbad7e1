#!/bin/bash
PCX_NVRTC_EXTRA="-DPCX_TWO_PASS=1 -DPCX_STAGGER_NS=19000" python tools/d3_timeline.py 2>&1 | tail -5
PCX_NVRTC_EXTRA="-DPCX_TWO_PASS=1" python tools/d3_timeline.py 2>&1 | tail -3
