#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools_trace.py > gpurun_out/trace9.txt 2>&1
timeout 600 python -m pytest tests -q --tb=short -p no:cacheprovider -m gpu -x > gpurun_out/tests9.log 2>&1
echo "tests exit $?" >> gpurun_out/summary9.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench9.json 2> gpurun_out/bench9.err
echo "bench exit $?" >> gpurun_out/summary9.txt
cat gpurun_out/summary9.txt; head -30 gpurun_out/trace9.txt; tail -5 gpurun_out/tests9.log | cut -c1-200; cat gpurun_out/bench9.json | cut -c1-300; grep -o '"kernel_ms": {[^}]*}' gpurun_out/bench9.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench9.json | cut -c1-200
