#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rulebook_gpu.py tests/test_conv_gpu.py tests/test_modules_gpu.py -q --tb=short -p no:cacheprovider -m gpu > gpurun_out/tests6.log 2>&1
echo "tests exit $?" >> gpurun_out/summary6.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench6.json 2> gpurun_out/bench6.err
echo "bench exit $?" >> gpurun_out/summary6.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gather" -s 10 -c 2 -o gpurun_out/prof_tc6 python bench.py --steps 2 --warmup 3 --graph 0 --cpu-sample 2000 > gpurun_out/ncu_full6.log 2>&1
echo "ncu_full exit $?" >> gpurun_out/summary6.txt
cat gpurun_out/summary6.txt; tail -15 gpurun_out/tests6.log; cat gpurun_out/bench6.json; tail -3 gpurun_out/bench6.err
