#!/bin/bash
# schedule sweep at N ranks: device time only (no e2e, no samples, no baseline), per-rank busy/wait from the trace
N=${1:-4}
O=gpurun_out/msweep$N
mkdir -p $O
for cfg in "2 4" "2 8" "1 2"; do
  set -- $cfg
  ( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N --steps 3 --warmup 2 \
      --schedule $1 --lookahead $2 --no-e2e --no-sample --no-cpu-baseline ) > $O/bench_s$1_l$2.json 2> $O/bench_s$1_l$2.err
  echo "== schedule $1 lookahead $2 rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_s$1_l$2.json").read().strip().splitlines()[-1])
    c=d.get("classes") or {}
    print("  ms", round(d["ms_per_step"],1), "GF", round(d["value"]), "parity", d["parity"]["factor_check_resident"], "waves", d["config"]["waves"], "rank0 busy/wait", round(c.get("busy_cta_s",0),1), round(c.get("wait_cta_s",0),1))
except Exception as e:
    print("  failed", e)
PY
done
