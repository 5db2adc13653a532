#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err
echo "bench_graph exit $?" >> gpurun_out/summary2.txt
timeout 600 python bench.py --steps 20 --warmup 5 --graph 0 > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err
echo "bench_eager exit $?" >> gpurun_out/summary2.txt
timeout 600 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench_ref exit $?" >> gpurun_out/summary2.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --graph 0 --cpu-sample 2000 > gpurun_out/ncu_launch.log 2>&1
echo "ncu_launches exit $?" >> gpurun_out/summary2.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:tc_ -s 12 -c 6 -o gpurun_out/prof_tc python bench.py --steps 2 --warmup 3 --graph 0 --cpu-sample 2000 > gpurun_out/ncu_full.log 2>&1
echo "ncu_full exit $?" >> gpurun_out/summary2.txt
cat gpurun_out/summary2.txt
cat gpurun_out/bench_graph.json
