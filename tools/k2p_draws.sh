# K2p: crossover draws before (product build) or after (-DMB2_POOL_EARLY_DRAWS=0) the wait for the row's buffers
# (the late-draws build: python -c "from mystic_b200 import build; build.build(extra_flags=('-DMB2_POOL_EARLY_DRAWS=0',), out='mystic_b200/_lib/libmystic_b200_latedraws.so')")
. tools/k2p_ab.sh.inc
qb "c2:MB2_TMA_POOL=1" | tee /tmp/first.txt
if grep -q FAILED /tmp/first.txt; then echo "pool kernel failed or hung: stop"; exit 1; fi
timeout 400 python -m pytest tests/test_gpu_pool.py tests/test_gpu_commit_modes.py -x -q 2>&1 | tail -4
L=$PWD/mystic_b200/_lib/libmystic_b200_latedraws.so
for rep in 1 2; do qb "c2:MB2_LIBRARY=$L" "c2:MB2_TMA_POOL=1" "c2:MB2_POOL_G=8,MB2_POOL_P=7,MB2_POOL_Q=12" "o2:MB2_TMA_POOL=1"; done
