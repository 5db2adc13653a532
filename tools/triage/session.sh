#!/bin/bash
mkdir -p gpurun_out
export SPX_SUBM_TABLE=grouped
timeout 900 python -m pytest tests/test_rulebook_gpu.py tests/test_modules_gpu.py -q --tb=short -p no:cacheprovider -m gpu -x > gpurun_out/tests28.log 2>&1
echo "grouped tests exit $?" > gpurun_out/summary28.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench28g.json 2> gpurun_out/bench28g.err
echo "grouped bench exit $?" >> gpurun_out/summary28.txt
timeout 200 python tools/triage/tools_probe_ablate.py > gpurun_out/probe_ablate28g.txt 2>&1
export SPX_SUBM_TABLE=flat
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench28f.json 2> gpurun_out/bench28f.err
echo "flat bench exit $?" >> gpurun_out/summary28.txt
cat gpurun_out/summary28.txt; tail -3 gpurun_out/tests28.log | cut -c1-300
for f in gpurun_out/bench28g.json gpurun_out/bench28f.json; do cut -c1-200 $f; grep -o '"kernel_ms": {[^}]*}' $f; grep -o '"e2e": {[^}]*}' $f | cut -c1-330; done
tail -9 gpurun_out/probe_ablate28g.txt
