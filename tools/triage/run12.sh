#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q --tb=short -p no:cacheprovider -m gpu -x > gpurun_out/tests12.log 2>&1
echo "tests exit $?" > gpurun_out/summary12.txt
timeout 120 python tools/triage/tools_trace_wgrad.py > gpurun_out/trace12w.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches12.csv python bench.py --graph 0 --steps 2 --warmup 1 > gpurun_out/ncu12.log 2>&1
echo "ncu exit $?" >> gpurun_out/summary12.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench12.json 2> gpurun_out/bench12.err
echo "bench exit $?" >> gpurun_out/summary12.txt
cat gpurun_out/summary12.txt; head -45 gpurun_out/trace12w.txt; tail -5 gpurun_out/tests12.log | cut -c1-200; cat gpurun_out/bench12.json | cut -c1-200; grep -o '"kernel_ms": {[^}]*}' gpurun_out/bench12.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench12.json | cut -c1-200
