#!/bin/bash
# A/B of build variants (run under gpurun): LIBS="path1;path2" ('' = the product library), WLS="c2 c5 x2"
IFS=";" read -ra LIBS_A <<< "${LIBS:-;}"
for rep in 1 2; do
for lib in "${LIBS_A[@]}"; do
for wl in ${WLS:-c2 x2 c5}; do
R=$(MB2_LIBRARY=$lib timeout 120 python bench.py --workload $wl --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['config']['burst_ms_per_step'], d['clocks']['sm_mhz'])")
echo "lib=${lib:-product} $wl -> $R"
done; done; done
