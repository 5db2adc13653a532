#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q --tb=short -p no:cacheprovider -m gpu > gpurun_out/tests8.log 2>&1
echo "tests exit $?" >> gpurun_out/summary8.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench8.json 2> gpurun_out/bench8.err
echo "bench exit $?" >> gpurun_out/summary8.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_|rs_" -s 40 -c 9 -o gpurun_out/prof_tc8 python bench.py --steps 2 --warmup 3 --graph 0 --cpu-sample 2000 > gpurun_out/ncu_full8.log 2>&1
echo "ncu_full exit $?" >> gpurun_out/summary8.txt
cat gpurun_out/summary8.txt; tail -8 gpurun_out/tests8.log | cut -c1-200; cat gpurun_out/bench8.json | cut -c1-1800; tail -3 gpurun_out/bench8.err | cut -c1-300
