# A/B of the pooled producer / consumer kernel K2p against K2a at BASELINE config 2 (run under gpurun)
qb() {  # one compact device-timed line, bounded
  for spec in "$@"; do
    envs=$(echo "${spec#*:}" | tr ',' ' ')
    env $envs timeout 90 python bench.py --workload ${spec%%:*} --also none --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$spec', 'ms %.4f burst %.4f frac %.4f burst_frac %.4f sm_mhz %s' % (d['ms_per_step'], d['details']['burst_ms_per_step'], d['roofline']['frac'], d['roofline'].get('frac_burst', 0), d['clocks']['sm_mhz']))
except Exception as e: print('$spec', 'FAILED', e)"
  done
}
qb "c2:MB2_TMA_POOL=1" "c2:MB2_TMA_POOL=1,MB2_POOL_GW=4" | tee /tmp/first.txt
if grep -q FAILED /tmp/first.txt; then echo "pool kernel failed or hung: stop"; exit 1; fi
MB2_TMA_POOL=1 timeout 200 python -m pytest tests/test_gpu_commit_modes.py tests/test_gpu_fullsize_parity.py -x -q -k "not config4 and not config5 and not config3" 2>&1 | tail -5
MB2_TMA_POOL=1 MB2_POOL_GW=4 timeout 200 python -m pytest tests/test_gpu_commit_modes.py tests/test_gpu_fullsize_parity.py -x -q -k "not config4 and not config5 and not config3" 2>&1 | tail -5
for rep in 1 2; do
qb "c2:MB2_TMA_POOL=0" "c2:MB2_TMA_POOL=1" "c2:MB2_POOL_GW=4" "c2:MB2_POOL_GW=4,MB2_POOL_P=8,MB2_POOL_Q=10" "c2:MB2_POOL_GW=4,MB2_POOL_P=9,MB2_POOL_Q=8" "c2:MB2_POOL_GW=4,MB2_POOL_P=10,MB2_POOL_Q=6" "c2:MB2_POOL_GW=4,MB2_POOL_G=3,MB2_POOL_P=9,MB2_POOL_Q=8" "c2:MB2_POOL_GW=2,MB2_POOL_G=7,MB2_POOL_P=8,MB2_POOL_Q=9"
done
