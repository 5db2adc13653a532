#!/bin/bash
# One GPU session = as much evidence as possible per box acquisition.  Edited between sessions;
# everything lands in gpurun_out/ (copied back by gpurun).  Usage on the box: bash tools/gpu_session.sh <tag>
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
echo "== session $TAG $(date -u +%H:%M:%S)"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40) > $O/${TAG}_tests.log
tail -4 $O/${TAG}_tests.log
(timeout 100 python tools/host_profile.py cfg2 50 > $O/${TAG}_hostprof_cfg2.log 2>&1); head -2 $O/${TAG}_hostprof_cfg2.log
(timeout 100 python tools/host_profile.py encoder 30 > $O/${TAG}_hostprof_enc.log 2>&1); head -2 $O/${TAG}_hostprof_enc.log
(timeout 400 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err); tail -c 400 $O/${TAG}_bench.err; head -c 600 $O/${TAG}_bench.json; echo
(timeout 120 python tools/ab_rulebook.py > $O/${TAG}_ab_rulebook.log 2>&1); cat $O/${TAG}_ab_rulebook.log | tail -20
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_ncu_launches.csv python bench.py --graph 0 --extras 0 --steps 2 --warmup 1 > $O/${TAG}_ncu_bench.log 2>&1); tail -2 $O/${TAG}_ncu_bench.log | cut -c1-200
echo "== done $(date -u +%H:%M:%S)"
