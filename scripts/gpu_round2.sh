#!/bin/bash
# Development pass: cr64 diagnosis at nb 256, parity tests of the kernels just changed, compact tile micro-benchmark, traced step at 100^3
O=gpurun_out/${1:-r02b}
mkdir -p $O
( time timeout 600 python scripts/debug_cr64.py --n 3000 ) > $O/debug_cr64_n3000.txt 2>&1
( time timeout 600 python scripts/debug_cr64.py --n 1200 --oracle 1 ) > $O/debug_cr64_n1200.txt 2>&1
( time timeout 1200 python -m pytest tests/test_gstrf_gpu.py tests/test_kernels_gpu.py tests/test_production_gpu.py -m gpu -x -q -p no:cacheprovider --timeout 900 -k "not multi" ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
( timeout 300 scripts/micro/ctile_bench ) > $O/ctile_bench.txt 2>&1
( time timeout 900 python scripts/trace_mega.py --size 100 ) > $O/trace100.txt 2>&1
( time PANGULU_SM100_LIB_DIR=$PWD/pangulu_b200/lib_prof CT_PROF=1 CT_PROF_MAX=8 timeout 900 python scripts/trace_mega.py --size 100 ) > $O/trace100_prof_small.txt 2>&1
( time PANGULU_SM100_LIB_DIR=$PWD/pangulu_b200/lib_prof CT_PROF=64 timeout 900 python scripts/trace_mega.py --size 100 ) > $O/trace100_prof_heavy.txt 2>&1
grep "phases" $O/trace100_prof_small.txt $O/trace100_prof_heavy.txt
tail -4 $O/pytest.log; cat $O/debug_cr64_n3000.txt $O/debug_cr64_n1200.txt | cut -c1-400; cat $O/ctile_bench.txt; head -12 $O/trace100.txt
