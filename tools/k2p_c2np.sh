# K2p at BASELINE config 2: one against two producer warps
. tools/k2p_ab.sh.inc
qb "c2:MB2_POOL_NPROD=2" | tee /tmp/first.txt
if grep -q FAILED /tmp/first.txt; then echo "pool kernel failed or hung: stop"; exit 1; fi
MB2_POOL_NPROD=2 timeout 300 python -m pytest tests/test_gpu_pool.py -x -q -k "full_grid or replayed" 2>&1 | tail -5
for rep in 1 2; do
qb "c2:MB2_POOL_NPROD=1" "c2:MB2_POOL_NPROD=2" "c2:MB2_POOL_NPROD=2,MB2_POOL_G=8,MB2_POOL_GW=2" "c2:MB2_POOL_NPROD=2,MB2_POOL_G=6" "c2:MB2_POOL_NPROD=2,MB2_POOL_P=8,MB2_POOL_Q=10"
done
