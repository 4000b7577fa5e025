# K2p on rows of 256 dimensions (BASELINE config 5's shape: one-warp consumer groups, two producer warps) against K2a
. tools/k2p_ab.sh.inc
qb "c5:MB2_POOL_SHORT=1" | tee /tmp/first.txt
if grep -q FAILED /tmp/first.txt; then echo "pool kernel failed or hung: stop"; exit 1; fi
MB2_POOL_SHORT=1 timeout 300 python -m pytest tests/test_gpu_commit_modes.py tests/test_gpu_fullsize_parity.py -x -q -k "D256 or config5 or config2_rosen" 2>&1 | tail -5
qb "c5:MB2_POOL_SHORT=0" "c5:MB2_POOL_SHORT=1" "c5:MB2_POOL_SHORT=1,MB2_POOL_G=18" "c5:MB2_POOL_SHORT=1,MB2_POOL_G=15" "c5:MB2_POOL_SHORT=1,MB2_POOL_NPROD=1,MB2_POOL_G=17" "c5:MB2_POOL_SHORT=1,MB2_POOL_P=24,MB2_POOL_Q=60" "c5:MB2_POOL_SHORT=1,MB2_POOL_P=36,MB2_POOL_Q=36" "c5:MB2_POOL_SHORT=0" "c5:MB2_POOL_SHORT=1" "c2:MB2_TMA_POOL=1"
