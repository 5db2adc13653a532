#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_modules_gpu.py -q --tb=short -p no:cacheprovider -m gpu -x > gpurun_out/tests19.log 2>&1
echo "tests exit $?" > gpurun_out/summary19.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench19.json 2> gpurun_out/bench19.err
echo "bench exit $?" >> gpurun_out/summary19.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"subm_|rs_|build_tile|tc_|wgrad_|tile_order" -c 700 --csv --log-file gpurun_out/launches19.csv python bench.py --graph 0 --steps 2 --warmup 1 > gpurun_out/ncu19.log 2>&1
echo "ncu exit $?" >> gpurun_out/summary19.txt
cat gpurun_out/summary19.txt; tail -5 gpurun_out/tests19.log | cut -c1-300; cat gpurun_out/bench19.json | cut -c1-200; grep -o '"kernel_ms": {[^}]*}' gpurun_out/bench19.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench19.json | cut -c1-200
python tools/launch_list.py gpurun_out/launches19.csv
