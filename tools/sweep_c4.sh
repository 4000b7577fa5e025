#!/bin/bash
# K2f geometry sweep on the c4 workload (run under gpurun): warps per CTA, stages per warp, pool slots (0 = what fits)
IFS=";" read -ra CFGS <<< "${SWEEP:-12 2 0;16 1 0;16 2 0;16 2 24;14 2 0;10 2 0;8 3 0;8 2 0;12 1 0;12 3 0}"
for cfg in "${CFGS[@]}"; do
set -- $cfg
R=$(MB2_RB_WARPS=$1 MB2_RB_STAGES=$2 MB2_RB_POOL=$3 timeout 100 python bench.py --workload ${WL:-c4} --no-cpu-baseline --sustain 0.2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['config']['block'], d['config']['smem_bytes'])")
echo "warps=$1 stages=$2 pool=$3 -> $R"
done
