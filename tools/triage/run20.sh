#!/bin/bash
mkdir -p gpurun_out
timeout 130 python tools/triage/tools_trace_wgrad.py > gpurun_out/trace20w.txt 2>&1
echo "trace exit $?" > gpurun_out/summary20.txt
cat gpurun_out/summary20.txt; head -12 gpurun_out/trace20w.txt; tail -4 gpurun_out/trace20w.txt
