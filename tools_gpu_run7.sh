#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_modules_gpu.py -v --tb=short -p no:cacheprovider -m gpu --timeout 100 > gpurun_out/tests7_modules.log 2>&1
echo "modules exit $?" >> gpurun_out/summary7.txt
timeout 600 python -m pytest tests/test_rulebook_gpu.py tests/test_conv_gpu.py -q --tb=short -p no:cacheprovider -m gpu --timeout 100 > gpurun_out/tests7.log 2>&1
echo "tests exit $?" >> gpurun_out/summary7.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench7.json 2> gpurun_out/bench7.err
echo "bench exit $?" >> gpurun_out/summary7.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_|rs_|subm_probe|build_tile" -s 40 -c 12 -o gpurun_out/prof_tc7 python bench.py --steps 2 --warmup 3 --graph 0 --cpu-sample 2000 > gpurun_out/ncu_full7.log 2>&1
echo "ncu_full exit $?" >> gpurun_out/summary7.txt
cat gpurun_out/summary7.txt; tail -25 gpurun_out/tests7_modules.log | cut -c1-200; tail -8 gpurun_out/tests7.log | cut -c1-200; cat gpurun_out/bench7.json | cut -c1-1500; tail -3 gpurun_out/bench7.err
