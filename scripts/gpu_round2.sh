#!/bin/bash
# Development pass: full parity suite, compact tile micro-benchmark, traced step at 100^3 (plain and with phase counters)
O=gpurun_out/${1:-r02b}
mkdir -p $O
( time timeout 600 python scripts/debug_cr64.py --n 1200 --oracle 1 ) > $O/debug_cr64_n1200.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 900 ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
( timeout 300 scripts/micro/ctile_bench ) > $O/ctile_bench.txt 2>&1
( time timeout 900 python scripts/trace_mega.py --size 100 ) > $O/trace100.txt 2>&1
( time PANGULU_SM100_LIB_DIR=$PWD/pangulu_b200/lib_prof CT_PROF=1 CT_PROF_MAX=8 timeout 900 python scripts/trace_mega.py --size 100 ) > $O/trace100_prof_small.txt 2>&1
( time PANGULU_SM100_LIB_DIR=$PWD/pangulu_b200/lib_prof CT_PROF=64 timeout 900 python scripts/trace_mega.py --size 100 ) > $O/trace100_prof_heavy.txt 2>&1
( time timeout 600 python scripts/quick_bench.py --case p3 --size 100 --nb 256 --steps 3 --profile 0 --mode 1 ) > $O/quick100.txt 2>&1
grep "phases" $O/trace100_prof_small.txt $O/trace100_prof_heavy.txt
tail -4 $O/pytest.log; head -3 $O/debug_cr64_n1200.txt | cut -c1-300; cat $O/ctile_bench.txt; head -12 $O/trace100.txt; grep "step\|factor" $O/quick100.txt
