#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/lstm_probe.py > gpurun_out/lstm_probe22.log 2>&1; echo "exit $?"; cat gpurun_out/lstm_probe22.log
