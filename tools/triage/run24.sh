#!/bin/bash
mkdir -p gpurun_out
for d in 0 16 32 48; do
  WG_DEBUG=$d timeout 130 python tools/triage/tools_trace_wgrad.py 2>&1 | tail -2 | sed "s/^/debug=$d: /"
done > gpurun_out/wgdump24.txt
cat gpurun_out/wgdump24.txt
