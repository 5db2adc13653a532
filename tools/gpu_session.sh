#!/bin/bash
# One GPU session = as much evidence as possible per box acquisition.  Edited between sessions;
# everything lands in gpurun_out/ (copied back by gpurun).  Usage on the box: bash tools/gpu_session.sh <tag>
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
echo "== session $TAG $(date -u +%H:%M:%S)"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -400) > $O/${TAG}_tests.log
tail -6 $O/${TAG}_tests.log | cut -c1-300
(timeout 400 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err); tail -c 300 $O/${TAG}_bench.err; head -c 400 $O/${TAG}_bench.json; echo
for W in second_encoder6_fp16 second_encoder6_fp16_b8; do for P in 2 0; do
(timeout 300 python bench.py --workload $W --pipeline $P --extras 0 --steps 10 > $O/${TAG}_bench_${W}_p$P.json 2>> $O/${TAG}_bench.err); python -c "import json; d=json.loads(open('$O/${TAG}_bench_${W}_p$P.json').read().strip().splitlines()[-1]); print('$W pipeline $P', round(d['value']/1e6,2), 'Mvox/s', round(d['ms_per_step'],4), 'ms; e2e', d['e2e'].get('ms_per_step'), d['e2e'].get('eager_variant'), d['e2e'].get('eager_prefetch_ms_per_step'), d['e2e'].get('eager_naive_ms_per_step'), 'launches', d['gpu_launches'])"
done; done
echo "== done $(date -u +%H:%M:%S)"
