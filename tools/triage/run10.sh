#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -q --tb=short -p no:cacheprovider -m gpu -x > gpurun_out/tests10.log 2>&1
echo "tests exit $?" > gpurun_out/summary10.txt
timeout 120 python tools/triage/tools_trace_wgrad.py > gpurun_out/trace10w.txt 2>&1
timeout 120 python tools/triage/tools_trace.py > gpurun_out/trace10f.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench10.json 2> gpurun_out/bench10.err
echo "bench exit $?" >> gpurun_out/summary10.txt
cat gpurun_out/summary10.txt; head -40 gpurun_out/trace10w.txt; tail -5 gpurun_out/tests10.log | cut -c1-200; cat gpurun_out/bench10.json | cut -c1-300; grep -o '"kernel_ms": {[^}]*}' gpurun_out/bench10.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench10.json | cut -c1-200
