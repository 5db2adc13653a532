#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_modules_gpu.py -q --tb=short -p no:cacheprovider -m gpu -x > gpurun_out/tests16.log 2>&1
echo "tests exit $?" > gpurun_out/summary16.txt
timeout 130 python tools/triage/tools_trace_wgrad.py > gpurun_out/trace16w.txt 2>&1
timeout 200 python tools/triage/tools_ablate.py > gpurun_out/ablate16.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench16.json 2> gpurun_out/bench16.err
echo "bench exit $?" >> gpurun_out/summary16.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 900 --csv --log-file gpurun_out/launches16_cfg4.csv python bench.py --graph 0 --steps 2 --warmup 1 --workload sparseconv3d_k3s2_c64_128_bf16_300k > gpurun_out/ncu16.log 2>&1
echo "ncu exit $?" >> gpurun_out/summary16.txt
cat gpurun_out/summary16.txt; head -30 gpurun_out/trace16w.txt; tail -3 gpurun_out/ablate16.txt; tail -5 gpurun_out/tests16.log | cut -c1-200; cat gpurun_out/bench16.json | cut -c1-200; grep -o '"kernel_ms": {[^}]*}' gpurun_out/bench16.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench16.json | cut -c1-200
