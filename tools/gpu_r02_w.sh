#!/bin/bash
python tools/d3_timeline.py 2>&1 | tail -24
