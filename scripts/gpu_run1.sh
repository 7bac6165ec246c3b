#!/bin/bash
# round-2 GPU call 1: baseline trace at the benchmark size + the whole GPU test suite (incl. the new production-parity tests)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r1_gpus.txt 2>&1
CT_PROF=1 timeout 600 python scripts/trace_mega.py --size 100 --nb 256 > gpurun_out/r1_trace100.txt 2>&1
echo "trace rc=$?" >> gpurun_out/r1_trace100.txt
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 1000 > gpurun_out/r1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r1_pytest.log
tail -5 gpurun_out/r1_trace100.txt; tail -30 gpurun_out/r1_pytest.log
