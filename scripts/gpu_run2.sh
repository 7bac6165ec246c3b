#!/bin/bash
# round-2 GPU call 2: tensor-loop microbenchmark, new compact tile routine (64-column tiles, runs): tests + trace at 100^3
mkdir -p gpurun_out
./scripts/micro/ctile_loop_bench > gpurun_out/r2_loop_bench.txt 2>&1
timeout 900 python -m pytest tests/test_gstrf_gpu.py tests/test_kernels_gpu.py tests/test_production_gpu.py tests/test_multigpu_gpu.py -m gpu -q -x -p no:cacheprovider --timeout 900 > gpurun_out/r2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
timeout 600 python scripts/trace_mega.py --size 100 --nb 256 > gpurun_out/r2_trace100.txt 2>&1
echo "trace rc=$?" >> gpurun_out/r2_trace100.txt
cat gpurun_out/r2_loop_bench.txt; tail -25 gpurun_out/r2_trace100.txt; tail -15 gpurun_out/r2_pytest.log
