#!/bin/bash
# One-GPU evidence run (under gpurun): default bench line + reference arm, the ncu launch list of the
# bench command, `ncu --set full` captures of the top kernels, the other workloads, smoke().
# Everything lands in gpurun_out/; tools/summarize_ncu.py and tools/ncu_lines.py turn the reports
# into the summaries tracked under profiles/.
set -x
python bench.py > gpurun_out/r1_bench_default.json 2> gpurun_out/r1_bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_reference.json 2> gpurun_out/r1_bench_reference.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_c2_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2_launches.log 2>&1
for w in ${CAPTURE:-}; do   # e.g. CAPTURE="c2:de_step_best1bin_tma c5:de_step_best1bin_tma c3:de_step_cheb_mma c4:de_step_rows_bulk x2:de_step_exp_tma"
  ncu --set full --clock-control none --import-source on -k regex:${w#*:} --launch-skip 3 -c 1 -f -o gpurun_out/prof_${w%%:*}_r1c python bench.py --workload ${w%%:*} --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/ncu_${w%%:*}_r1c.log 2>&1
done
for w in c3 c4 c5 f2 x2; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/r1_bench_${w}.json 2> gpurun_out/r1_bench_${w}.err; done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1c.log 2>&1; tail -2 gpurun_out/smoke_r1c.log
tail -c 600 gpurun_out/r1_bench_default.json; tail -c 400 gpurun_out/r1_bench_reference.json
for w in c3 c4 c5 f2 x2; do python -c "
import json; d=json.load(open('gpurun_out/r1_bench_$w.json')); print('$w', d['ms_per_step'], '%.4g'%d['value'], d['roofline']['frac'], 'e2e %.4g'%d['e2e']['value'])"; done
