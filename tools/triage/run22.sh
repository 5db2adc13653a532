#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q --tb=short -p no:cacheprovider -m gpu -x > gpurun_out/tests22.log 2>&1
echo "tests exit $?" > gpurun_out/summary22.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke22.log 2>&1
echo "smoke exit $?" >> gpurun_out/summary22.txt
timeout 400 python bench.py > gpurun_out/bench22.json 2> gpurun_out/bench22.err
echo "bench exit $?" >> gpurun_out/summary22.txt
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench22_ref.json 2> gpurun_out/bench22_ref.err
echo "bench ref exit $?" >> gpurun_out/summary22.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"subm_|rs_|build_tile|tc_|wgrad_|tile_order" -c 700 --csv --log-file gpurun_out/launches22.csv python bench.py --graph 0 --steps 2 --warmup 1 > gpurun_out/ncu22.log 2>&1
echo "ncu list exit $?" >> gpurun_out/summary22.txt
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"tc_wgrad|tc_gather|wgrad_reduce" --launch-skip 30 -c 4 -o gpurun_out/prof_tc22 -f python bench.py --graph 0 --steps 3 --warmup 1 > gpurun_out/ncu22_full.log 2>&1
echo "ncu full exit $?" >> gpurun_out/summary22.txt
cat gpurun_out/summary22.txt; tail -3 gpurun_out/tests22.log | cut -c1-300; tail -2 gpurun_out/smoke22.log; cat gpurun_out/bench22.json | cut -c1-250; grep -o '"kernel_ms": {[^}]*}' gpurun_out/bench22.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench22.json | cut -c1-200; cat gpurun_out/bench22_ref.json | cut -c1-400
python tools/launch_list.py gpurun_out/launches22.csv
