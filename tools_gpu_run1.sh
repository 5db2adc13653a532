#!/bin/bash
# first GPU session: rulebook parity, SIMT conv parity, then tcgen05 conv parity
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_rulebook_gpu.py -q --tb=short -p no:cacheprovider -m gpu > gpurun_out/rb.log 2>&1
echo "rulebook exit $?" >> gpurun_out/summary.txt
SPX_FORCE_SIMT=1 timeout 1200 python -m pytest tests/test_conv_gpu.py -q --tb=short -p no:cacheprovider -m gpu > gpurun_out/conv_simt.log 2>&1
echo "conv_simt exit $?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_conv_gpu.py -q --tb=short -p no:cacheprovider -m gpu -k "f16 and C64K64 and subm" > gpurun_out/conv_tc_first.log 2>&1
echo "conv_tc_first exit $?" >> gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_conv_gpu.py -q --tb=short -p no:cacheprovider -m gpu > gpurun_out/conv_tc.log 2>&1
echo "conv_tc exit $?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/rb.log gpurun_out/conv_simt.log gpurun_out/conv_tc_first.log gpurun_out/conv_tc.log gpurun_out/smoke.log
