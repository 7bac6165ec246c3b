#!/bin/bash
# multi-GPU call: parity tests on the GPUs of the box, then the bench at N = number of GPUs
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi${N}_gpus.txt
timeout 1200 python -m pytest tests/test_multigpu_gpu.py tests/test_production_gpu.py -m gpu -q -x -p no:cacheprovider --timeout 900 -k "multi" > gpurun_out/multi${N}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/multi${N}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/multi${N}_bench.json 2> gpurun_out/multi${N}_bench.err
echo "bench rc=$?" >> gpurun_out/multi${N}_bench.err
tail -4 gpurun_out/multi${N}_pytest.log; tail -3 gpurun_out/multi${N}_bench.err; tail -c 1500 gpurun_out/multi${N}_bench.json
