#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/timeline.py --graph --out gpurun_out/timeline15 > gpurun_out/timeline15.log 2>&1; echo "timeline exit $?"; head -n 60 gpurun_out/timeline15.log
