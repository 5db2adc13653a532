#!/bin/bash
# multi-GPU session: bash tools/gpu_session_multi.sh <tag> <N>
TAG=${1:-r2n}; N=${2:-2}
O=gpurun_out
mkdir -p $O
echo "== multi session $TAG N=$N $(date -u +%H:%M:%S)"; nvidia-smi --query-gpu=name --format=csv,noheader | head -1
run() {  # name, extra args
  (timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 $2 > $O/${TAG}_$1.json 2> $O/${TAG}_$1.err)
  python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_$1.json").read().strip().splitlines()[-1])
    print("$1", "value", round(d["value"]/1e6,1), "M/s ms", round(d["ms_per_step"],4), "serial", d.get("serial_ms_per_step"), "e2e", round(d["e2e"]["ms_per_step"],4), "eager", round(d["e2e"]["eager_ms_per_step"],4))
    for k,w in d.get("workloads",{}).items():
        print("   ", k, w.get("error") or (round(w["value"]/1e6,1), round(w["ms_per_step"],4)))
except Exception as e:
    print("$1 FAILED", e); print(open("$O/${TAG}_$1.err").read()[-500:])
PY
}
run bench ""
(timeout 120 python bench.py --gpus 1 --extras 0 > $O/${TAG}_bench_n1.json 2>> $O/${TAG}_bench.err); python -c "import json; d=json.loads(open('$O/${TAG}_bench_n1.json').read().strip().splitlines()[-1]); print('N=1 same box', d['value']/1e6, d['ms_per_step'])"
run bench_fused "--allreduce fused --extras 0"
if [ "${PEER_CHECK:-1}" = "1" ]; then (timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/peer_check.py > $O/${TAG}_peer_check.log 2>&1); grep -v Warning $O/${TAG}_peer_check.log | tail -2 | cut -c1-400; fi
echo "== done $(date -u +%H:%M:%S)"
