#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/triage/gather4_probe > gpurun_out/gather4_11.txt 2>&1
echo "probe exit $?" > gpurun_out/summary11.txt
timeout 600 python -m pytest tests/test_rulebook_gpu.py -q --tb=short -p no:cacheprovider -m gpu -x > gpurun_out/tests11.log 2>&1
echo "tests exit $?" >> gpurun_out/summary11.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches11.csv python bench.py --graph 0 --steps 2 --warmup 1 > gpurun_out/ncu11.log 2>&1
echo "ncu exit $?" >> gpurun_out/summary11.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench11.json 2> gpurun_out/bench11.err
echo "bench exit $?" >> gpurun_out/summary11.txt
cat gpurun_out/summary11.txt; cat gpurun_out/gather4_11.txt; tail -5 gpurun_out/tests11.log | cut -c1-200; cat gpurun_out/bench11.json | cut -c1-200; grep -o '"kernel_ms": {[^}]*}' gpurun_out/bench11.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench11.json | cut -c1-200
