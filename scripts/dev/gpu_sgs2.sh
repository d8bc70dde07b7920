#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_sgs.py -q -x 2>&1 | tail -30 | tee gpurun_out/pytest_sgs.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_callers.json 2> gpurun_out/bench_callers.err; tail -3 gpurun_out/bench_callers.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_callers.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'])
print(json.dumps(d['callers'], indent=1))
PY
