#!/bin/bash
# K2e ring-geometry sweep on the x2 workload (run under gpurun): groups, warps per group, stages, cap
IFS=";" read -ra CFGS <<< "${SWEEP:-0 0 0 0;15 2 1 0;10 2 2 0;20 1 1 0;10 1 2 0;7 2 3 32;11 2 2 32}"
for cfg in "${CFGS[@]}"; do
set -- $cfg
R=$(MB2_XTMA_GROUPS=$1 MB2_XTMA_GWARPS=$2 MB2_XTMA_STAGES=$3 MB2_XTMA_CAP=$4 timeout 100 python bench.py --workload ${WL:-x2} --no-cpu-baseline --sustain 0.2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['config']['block'], d['config']['smem_bytes'])")
echo "g=$1 gw=$2 st=$3 cap=$4 -> $R"
done
