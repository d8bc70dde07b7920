#!/bin/bash
# development sweep of the bin occupancy / initial-margin tunables (search kernel)
for cfg in ${CFGS:-C2}; do
for div in 16 24 32 48; do for mf in 0.9 1.0; do
  echo -n "$cfg div=$div mf=$mf: "
  GSK_BIN_OCC_DIV=$div GSK_MARGIN_FACTOR=$mf timeout 300 python bench.py --config $cfg --steps 3 --warmup 3 --targets ${TGT:-0} --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('search %.3f solve %.3f step %.3f' % (d['phases_ms']['search'], d['phases_ms']['solve'], d['ms_per_step']))"
done; done; done
