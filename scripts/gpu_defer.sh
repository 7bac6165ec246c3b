#!/bin/bash
# task deferral in the persistent executor: one GPU (quick bench, defer 0/2/4) and N ranks (bench device time + distributed parity, defer 0/2/4)
N=${1:-2}
O=gpurun_out/defer$N
mkdir -p $O
for d in 0 2 4; do
  ( timeout 300 python scripts/quick_bench.py --case p3 --size 100 --nb 256 --steps 2 --profile 0 --mode 1 --opt defer=$d --opt spin_timeout_ms=10000 ) > $O/one_gpu_defer$d.txt 2>&1
  echo "== 1 GPU defer $d: $(grep 'step 1' $O/one_gpu_defer$d.txt) $(grep 'factor' $O/one_gpu_defer$d.txt) $(grep -c ERROR $O/one_gpu_defer$d.txt) errors"
done
for d in 0 2 4; do
  ( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus $N --steps 3 --warmup 2 \
      --defer $d --no-e2e --no-sample --no-cpu-baseline ) > $O/bench_n${N}_defer$d.json 2> $O/bench_n${N}_defer$d.err
  echo "== $N GPUs defer $d rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_n${N}_defer$d.json").read().strip().splitlines()[-1])
    c=d.get("classes") or {}
    print("  ms", round(d["ms_per_step"],1), "GF", round(d["value"]), "parity", d["parity"]["factor_check_resident"], "rank0 busy/wait", round(c.get("busy_cta_s",0),1), round(c.get("wait_cta_s",0),1))
except Exception as e:
    print("  failed", e)
PY
done
