#!/bin/bash
# compact device-timed lines: tools/quick_bench.sh c2 c5 "c4:MB2_RB_LDGSTS=0" ...   (workload[:ENV=VALUE[,ENV=VALUE]])
for spec in "$@"; do
  w=${spec%%:*}; envs=""
  if [[ "$spec" == *:* ]]; then envs=$(echo "${spec#*:}" | tr ',' ' '); fi
  env $envs python bench.py --workload $w --also none --steps ${STEPS:-20} --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$spec', 'ms %.4f burst %.4f frac %.4f burst_frac %.4f sm_mhz %s e2e %.4g' % (d['ms_per_step'], d['details']['burst_ms_per_step'], d['roofline']['frac'], d['roofline'].get('frac_burst', 0), d['clocks']['sm_mhz'], d['e2e']['value'] or 0))
except Exception as e: print('$spec', 'FAILED', e)"
done
