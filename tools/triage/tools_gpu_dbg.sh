#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 100 python -X faulthandler -m pytest "tests/test_modules_gpu.py" -v --tb=short -p no:cacheprovider -m gpu -o faulthandler_timeout=30 -x > gpurun_out/dbg_modules_$i.log 2>&1
echo "dbg$i exit $?" >> gpurun_out/summary_dbg.txt
done
OMP_NUM_THREADS=8 timeout 100 python -X faulthandler -m pytest "tests/test_modules_gpu.py" -v --tb=short -p no:cacheprovider -m gpu -o faulthandler_timeout=30 -x > gpurun_out/dbg_modules_omp8.log 2>&1
echo "dbg_omp8 exit $?" >> gpurun_out/summary_dbg.txt
cat gpurun_out/summary_dbg.txt
grep -B2 -A40 "Timeout\|Thread 0x\|most recent call" gpurun_out/dbg_modules_1.log | head -120 | cut -c1-200
tail -5 gpurun_out/dbg_modules_2.log gpurun_out/dbg_modules_omp8.log | cut -c1-200
