# parity of K2p (tests/test_gpu_pool.py + the commit-mode and full-size cases) and a short A/B against K2a (under gpurun)
timeout 400 python -m pytest tests/test_gpu_pool.py tests/test_gpu_commit_modes.py tests/test_gpu_fullsize_parity.py -q -k "not config4 and not config5 and not config3" 2>&1 | tail -15
. tools/k2p_ab.sh.inc
for rep in 1; do qb "c2:MB2_TMA_POOL=0" "c2:MB2_TMA_POOL=1"; done
