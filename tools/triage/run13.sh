#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_modules_gpu.py -q --tb=short -p no:cacheprovider -m gpu -x > gpurun_out/tests13.log 2>&1
echo "tests exit $?" > gpurun_out/summary13.txt
timeout 130 python tools/triage/tools_trace_wgrad.py > gpurun_out/trace13w.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/launches13.csv python bench.py --graph 0 --steps 2 --warmup 1 > gpurun_out/ncu13.log 2>&1
echo "ncu exit $?" >> gpurun_out/summary13.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench13.json 2> gpurun_out/bench13.err
echo "bench exit $?" >> gpurun_out/summary13.txt
cat gpurun_out/summary13.txt; head -45 gpurun_out/trace13w.txt; tail -5 gpurun_out/tests13.log | cut -c1-200; cat gpurun_out/bench13.json | cut -c1-200; grep -o '"kernel_ms": {[^}]*}' gpurun_out/bench13.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench13.json | cut -c1-200
