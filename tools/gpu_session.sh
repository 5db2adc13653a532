#!/bin/bash
# One GPU session = as much evidence as possible per box acquisition.  Edited between sessions;
# everything lands in gpurun_out/ (copied back by gpurun).  Usage on the box: bash tools/gpu_session.sh <tag>
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
echo "== session $TAG $(date -u +%H:%M:%S)"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
(timeout 170 python -m pytest tests -m gpu -q -x 2>&1 | tail -400) > $O/${TAG}_tests.log
tail -6 $O/${TAG}_tests.log | cut -c1-300
(timeout 200 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err); tail -c 300 $O/${TAG}_bench.err; head -c 400 $O/${TAG}_bench.json; echo
(timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -2)
echo "== done $(date -u +%H:%M:%S)"
