#!/bin/bash
# validate the overlapped factor download (host-buffer gstrf): parity tests that use mode 2, then the default bench line
O=gpurun_out/${1:-r02g}
mkdir -p $O
( time timeout 900 python -m pytest tests/test_production_gpu.py tests/test_gstrf_gpu.py tests/test_abi.py -m gpu -x -q -p no:cacheprovider --timeout 600 -k "not multi" ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
( time timeout 900 python bench.py --no-cpu-baseline ) > $O/bench_default.json 2> $O/bench_default.err
echo "bench rc=$?" >> $O/bench_default.err
tail -4 $O/pytest.log; tail -3 $O/bench_default.err; python - <<PY
import json
d=json.loads(open("$O/bench_default.json").read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"],"parity",d["parity"])
PY
