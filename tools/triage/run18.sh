#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q --tb=short -p no:cacheprovider -m gpu -x > gpurun_out/tests18.log 2>&1
echo "tests exit $?" > gpurun_out/summary18.txt
timeout 130 python tools/triage/tools_cta_spans.py > gpurun_out/spans18.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench18.json 2> gpurun_out/bench18.err
echo "bench exit $?" >> gpurun_out/summary18.txt
timeout 300 python bench.py --steps 10 --warmup 3 --workload sparseconv3d_k3s2_c64_128_bf16_300k > gpurun_out/bench18_cfg4.json 2> gpurun_out/bench18_cfg4.err
echo "bench cfg4 exit $?" >> gpurun_out/summary18.txt
cat gpurun_out/summary18.txt; head -14 gpurun_out/spans18.txt; tail -8 gpurun_out/tests18.log | cut -c1-300; cat gpurun_out/bench18.json | cut -c1-200; grep -o '"kernel_ms": {[^}]*}' gpurun_out/bench18.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench18.json | cut -c1-200
cat gpurun_out/bench18_cfg4.json | cut -c1-200; grep -o '"kernel_ms": {[^}]*}' gpurun_out/bench18_cfg4.json; tail -3 gpurun_out/bench18_cfg4.err
