O=gpurun_out; mkdir -p $O
for V in "nccl:--allreduce nccl" "local:--allreduce fused-local" "nofinish:--allreduce fused-local --peer-triage 1" "nopublish:--allreduce fused-local --peer-triage 2" "unforked:--allreduce fused-local --peer-triage 4" "neither:--allreduce fused-local --peer-triage 3"; do
  N=${V%%:*}; A=${V#*:}
  (timeout 200 python bench.py --extras 0 $A > $O/tri_$N.json 2> $O/tri_$N.err)
  python - <<PY
import json
try:
    d=json.loads(open("$O/tri_$N.json").read().strip().splitlines()[-1])
    print("$N", "value ms", round(d["ms_per_step"],4), "serial", round(d.get("serial_ms_per_step") or 0,4), "e2e", round(d["e2e"]["ms_per_step"],4), {k:v for k,v in d["kernel_ms"].items() if "gemm" in k})
except Exception as e:
    print("$N FAILED", e); print(open("$O/tri_$N.err").read()[-600:])
PY
done
