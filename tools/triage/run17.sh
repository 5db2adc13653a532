#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rulebook_gpu.py -q --tb=short -p no:cacheprovider -m gpu -x > gpurun_out/tests17.log 2>&1
echo "tests exit $?" > gpurun_out/summary17.txt
timeout 130 python tools/triage/tools_cta_spans.py > gpurun_out/spans17.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench17.json 2> gpurun_out/bench17.err
echo "bench exit $?" >> gpurun_out/summary17.txt
cat gpurun_out/summary17.txt; cat gpurun_out/spans17.txt | tail -40; tail -5 gpurun_out/tests17.log | cut -c1-200; cat gpurun_out/bench17.json | cut -c1-200; grep -o '"kernel_ms": {[^}]*}' gpurun_out/bench17.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench17.json | cut -c1-200
