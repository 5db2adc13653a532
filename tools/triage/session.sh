#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py -q --tb=short -p no:cacheprovider -m gpu -x -k "full_size or scheduler" > gpurun_out/tests30.log 2>&1
echo "tests exit $?" > gpurun_out/summary30.txt
timeout 300 python bench.py --steps 10 --warmup 3 --workload sparseconv3d_k3s2_c64_128_bf16_300k > gpurun_out/bench30_cfg4.json 2> gpurun_out/bench30_cfg4.err
echo "cfg4 bench exit $?" >> gpurun_out/summary30.txt
cat gpurun_out/summary30.txt; tail -15 gpurun_out/tests30.log | cut -c1-300; cut -c1-250 gpurun_out/bench30_cfg4.json; grep -o '"kernel_ms": {[^}]*}' gpurun_out/bench30_cfg4.json
