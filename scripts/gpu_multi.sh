#!/bin/bash
# multi-GPU call: parity tests on the GPUs of the box, then the bench at N = number of GPUs (and at N/2)
N=${1:-2}
O=gpurun_out/multi$N
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
( time timeout 1200 python -m pytest tests/test_multigpu_gpu.py tests/test_production_gpu.py -m gpu -q -x -p no:cacheprovider --timeout 900 -k "multi" ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
for n in $N $((N/2)); do
  [ $n -lt 2 ] && continue
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $n --steps 3 --warmup 3 ) > $O/bench_n$n.json 2> $O/bench_n$n.err
  echo "bench rc=$?" >> $O/bench_n$n.err
done
tail -4 $O/pytest.log; for n in $N $((N/2)); do tail -3 $O/bench_n$n.err; tail -1 $O/bench_n$n.json | head -c 700; echo; done
