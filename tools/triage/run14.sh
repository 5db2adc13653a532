#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rulebook_gpu.py tests/test_conv_gpu.py -q --tb=short -p no:cacheprovider -m gpu -x > gpurun_out/tests14.log 2>&1
echo "tests exit $?" > gpurun_out/summary14.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"subm_|rs_|build_tile|tc_|wgrad_" -c 200 --csv --log-file gpurun_out/launches14.csv python bench.py --graph 0 --steps 2 --warmup 1 > gpurun_out/ncu14.log 2>&1
echo "ncu exit $?" >> gpurun_out/summary14.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench14.json 2> gpurun_out/bench14.err
echo "bench exit $?" >> gpurun_out/summary14.txt
timeout 300 python bench.py --steps 10 --warmup 3 --workload sparseconv3d_k3s2_c64_128_bf16_300k > gpurun_out/bench14_cfg4.json 2> gpurun_out/bench14_cfg4.err
echo "bench cfg4 exit $?" >> gpurun_out/summary14.txt
cat gpurun_out/summary14.txt; tail -5 gpurun_out/tests14.log | cut -c1-200; cat gpurun_out/bench14.json | cut -c1-200; grep -o '"kernel_ms": {[^}]*}' gpurun_out/bench14.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench14.json | cut -c1-200
cat gpurun_out/bench14_cfg4.json | cut -c1-300; grep -o '"kernel_ms": {[^}]*}' gpurun_out/bench14_cfg4.json; tail -3 gpurun_out/bench14_cfg4.err
python tools/launch_list.py gpurun_out/launches14.csv
