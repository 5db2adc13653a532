#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_modules_gpu.py -q --tb=short -p no:cacheprovider -m gpu -x > gpurun_out/tests23.log 2>&1
echo "tests exit $?" > gpurun_out/summary23.txt
timeout 130 python tools/triage/tools_trace_wgrad.py > gpurun_out/trace23w.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench23.json 2> gpurun_out/bench23.err
echo "bench exit $?" >> gpurun_out/summary23.txt
cat gpurun_out/summary23.txt; head -3 gpurun_out/trace23w.txt; tail -2 gpurun_out/trace23w.txt; tail -5 gpurun_out/tests23.log | cut -c1-300; cat gpurun_out/bench23.json | cut -c1-200; grep -o '"kernel_ms": {[^}]*}' gpurun_out/bench23.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench23.json | cut -c1-200
