# Triage of the fused weight-gradient exchange on ONE GPU (world-of-one group): which part of the captured backward costs what.
# Results: profiles/r02_triage_*_n1.json, table in profiles/README.md.  The "nopublish" variant of the first run spun into the
# finish timeout (a finish without a publish waits for nothing) and is not in the list any more.  The committed numbers were
# taken when the captured data-parallel backward still ran the two gradients one after the other ("sequential"); it now keeps
# them as parallel branches (pytorch/ops.py), so a re-run measures the shipped shape.
O=gpurun_out; mkdir -p $O
for V in "nccl:--allreduce nccl" "local:--allreduce fused-local" "nofinish:--allreduce fused-local --peer-triage 1" "sequential:--allreduce fused-local --peer-triage 3"; do
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
