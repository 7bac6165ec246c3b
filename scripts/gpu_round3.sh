#!/bin/bash
# Evidence pass: reference CUDA platform as on-box comparator, nb sweep, default bench line (with the device-resident e2e flavour)
O=gpurun_out/${1:-r02e}
mkdir -p $O
( time timeout 900 python scripts/ref_cuda_compare.py 32 48 ) > $O/ref_cuda_compare.txt 2> $O/ref_cuda_compare.err
for nb in 64 128 256 512; do
  ( time timeout 300 python scripts/quick_bench.py --case p3 --size 64 --nb $nb --steps 3 --profile 0 --mode 1 ) > $O/nb_sweep_64_nb$nb.txt 2>&1
done
( time timeout 900 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err
echo "bench rc=$?" >> $O/bench_default.err
cat $O/ref_cuda_compare.txt; tail -3 $O/ref_cuda_compare.err; grep -H "step 2\|factor\|ERROR" $O/nb_sweep_64_nb*.txt | cut -c1-200; head -c 600 $O/bench_default.json; tail -3 $O/bench_default.err
