#!/bin/bash
# Round-2 ncu captures (one GPU, under gpurun): `--set full` + source of the fused kernel of each workload named in
# $CAPTURE ("c5:de_step_best1bin_tma c4:de_step_rows_bulk ..."), after the same command has run clean without ncu.
TAG=${TAG:-r2}
for w in ${CAPTURE:-}; do
  python bench.py --workload ${w%%:*} --also none --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/plain_${w%%:*}_${TAG}.log 2>&1 || { echo "plain run of ${w%%:*} failed"; continue; }
  ncu --set full --clock-control none --import-source on -k regex:${w#*:} --launch-skip 3 -c 1 -f -o gpurun_out/prof_${w%%:*}_${TAG} python bench.py --workload ${w%%:*} --also none --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/ncu_${w%%:*}_${TAG}.log 2>&1
  ls -la gpurun_out/prof_${w%%:*}_${TAG}.ncu-rep
done
