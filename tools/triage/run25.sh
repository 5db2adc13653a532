#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus25.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench25_n2.json 2> gpurun_out/bench25_n2.err
echo "bench n2 exit $?" > gpurun_out/summary25.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/bench25_n2_ref.json 2> gpurun_out/bench25_n2_ref.err
echo "bench n2 ref exit $?" >> gpurun_out/summary25.txt
cat gpurun_out/summary25.txt; cat gpurun_out/gpus25.txt; tail -3 gpurun_out/bench25_n2.err; cat gpurun_out/bench25_n2.json | cut -c1-400; grep -o '"e2e": {[^}]*}' gpurun_out/bench25_n2.json | cut -c1-300; cat gpurun_out/bench25_n2_ref.json | cut -c1-300
