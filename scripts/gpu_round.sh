#!/bin/bash
# One-GPU evidence pass: parity tests, the default bench line, ncu launch list + DRAM traffic of the persistent kernel at the
# benchmark size, bench lines of the other BASELINE configurations.  Every step logs to gpurun_out/ separately.
O=gpurun_out/${1:-round}
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 900 ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
( time timeout 900 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err
echo "bench rc=$?" >> $O/bench_default.err
# DRAM traffic of ONE launch of the persistent kernel at 100^3 (counter-only pass, no replay of the 1.2 s kernel needed)
( time timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:mega_kernel -c 1 --csv \
    --log-file $O/ncu_dram_mega100.csv python scripts/quick_bench.py --case p3 --size 100 --nb 256 --steps 1 --profile 0 --mode 1 ) > $O/ncu_dram.log 2>&1
# launch list of the bench command (shares of the step, not absolute times)
( time timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sample --no-classes --no-e2e ) > $O/launches.log 2>&1
# one full-section capture of the dominant routine as a stand-alone kernel (per-wave executor), 64^3: a launch late in the factorisation
( time timeout 600 ncu --set full --clock-control none --import-source on -k regex:ssssm_ctile_kernel --launch-skip 150 --launch-count 1 -f -o $O/ncu_ctile_full \
    python scripts/quick_bench.py --case p3 --size 64 --nb 256 --steps 1 --profile 0 --mode 1 --opt executor=0 ) > $O/ncu_ctile_full.log 2>&1
for w in laplace2d_1000 stencil27_64 stencil27_64_r32 circuit_cr64; do
  ( time timeout 900 python bench.py --workload $w ) > $O/bench_$w.json 2> $O/bench_$w.err
  echo "bench rc=$?" >> $O/bench_$w.err
done
tail -3 $O/pytest.log; for f in $O/bench_*.json; do echo "== $f"; head -c 400 $f; echo; done; tail -5 $O/ncu_dram_mega100.csv
