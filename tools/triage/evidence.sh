#!/bin/bash
# One GPU session that produces everything profiles/ cites (run under gpurun from the repo root):
#   tests, smoke, both bench arms, the ncu launch list and one --set full capture of the GEMM kernels.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q --tb=short -p no:cacheprovider -m gpu -x > gpurun_out/tests.log 2>&1
echo "tests exit $?" > gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/summary.txt
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench ref exit $?" >> gpurun_out/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"subm_|rs_|build_tile|tc_|wgrad_|tile_order" -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --graph 0 --steps 2 --warmup 1 > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?" >> gpurun_out/summary.txt
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"tc_wgrad|tc_gather|wgrad_reduce" --launch-skip 30 -c 4 -o gpurun_out/prof_tc -f python bench.py --graph 0 --steps 3 --warmup 1 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/tests.log | cut -c1-300; tail -2 gpurun_out/smoke.log
cut -c1-250 gpurun_out/bench.json; cut -c1-300 gpurun_out/bench_ref.json
python tools/launch_list.py gpurun_out/launches.csv
