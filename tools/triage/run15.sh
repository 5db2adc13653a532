#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rulebook_gpu.py tests/test_modules_gpu.py -q --tb=short -p no:cacheprovider -m gpu -x > gpurun_out/tests15.log 2>&1
echo "tests exit $?" > gpurun_out/summary15.txt
timeout 400 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"subm_probe|build_tile|rs_scatter|subm_insert" --launch-skip 40 -c 6 -o gpurun_out/prof_rb15 -f python bench.py --graph 0 --steps 3 --warmup 1 > gpurun_out/ncu15.log 2>&1
echo "ncu exit $?" >> gpurun_out/summary15.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench15.json 2> gpurun_out/bench15.err
echo "bench exit $?" >> gpurun_out/summary15.txt
cat gpurun_out/summary15.txt; tail -5 gpurun_out/tests15.log | cut -c1-200; cat gpurun_out/bench15.json | cut -c1-200; grep -o '"kernel_ms": {[^}]*}' gpurun_out/bench15.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench15.json | cut -c1-200
