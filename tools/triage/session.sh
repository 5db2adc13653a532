#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q --tb=short -p no:cacheprovider -m gpu -x > gpurun_out/tests27.log 2>&1
echo "tests exit $?" > gpurun_out/summary27.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench27.json 2> gpurun_out/bench27.err
echo "bench exit $?" >> gpurun_out/summary27.txt
cat gpurun_out/summary27.txt; tail -3 gpurun_out/tests27.log | cut -c1-300; cut -c1-200 gpurun_out/bench27.json; grep -o '"kernel_ms": {[^}]*}' gpurun_out/bench27.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench27.json | cut -c1-200; tail -3 gpurun_out/bench27.err
