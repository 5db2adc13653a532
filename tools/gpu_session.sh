#!/bin/bash
# One GPU session = as much evidence as possible per box acquisition.  Edited between sessions;
# everything lands in gpurun_out/ (copied back by gpurun).  Usage on the box: bash tools/gpu_session.sh <tag>
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
echo "== session $TAG $(date -u +%H:%M:%S)"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -400) > $O/${TAG}_tests.log
tail -6 $O/${TAG}_tests.log | cut -c1-300
(timeout 100 python tools/host_profile.py cfg2 50 > $O/${TAG}_hostprof_cfg2.log 2>&1); head -1 $O/${TAG}_hostprof_cfg2.log
(timeout 100 python tools/host_profile.py encoder 30 > $O/${TAG}_hostprof_enc.log 2>&1); head -1 $O/${TAG}_hostprof_enc.log
(timeout 400 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err); tail -c 300 $O/${TAG}_bench.err; head -c 400 $O/${TAG}_bench.json; echo
(timeout 300 python bench.py --workload second_encoder6_fp16_b8 --extras 0 --steps 10 > $O/${TAG}_bench_enc_b8.json 2>> $O/${TAG}_bench.err); python -c "import json; d=json.loads(open('$O/${TAG}_bench_enc_b8.json').read().strip().splitlines()[-1]); print('encoder b8', d['value']/1e6, d['ms_per_step'], d.get('kernel_ms_summary'))"
(timeout 200 python bench.py --impl reference --steps 20 --warmup 3 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err); head -c 200 $O/${TAG}_bench_reference.json; echo
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_ncu_launches.csv python bench.py --graph 0 --extras 0 --steps 2 --warmup 1 > $O/${TAG}_ncu_bench.log 2>&1); tail -1 $O/${TAG}_ncu_bench.log | cut -c1-120
echo "== done $(date -u +%H:%M:%S)"
