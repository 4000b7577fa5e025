#!/bin/bash
# One-GPU evidence run of round 2 (under gpurun): the driver's bench command and the reference arm, the other
# workloads (default commit mode = auto, and the BASELINE ones again with the double buffer forced), the ncu launch
# list of the bench command, `ncu --set full` captures of the fused kernels committing in place, smoke().
# Everything lands in gpurun_out/; tools/summarize_ncu.py and tools/ncu_lines.py turn the reports into profiles/.
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
: > gpurun_out/r2_bench_1gpu.jsonl
for w in c2 c4 c5 c3 c3h x2 o2 ox2 g128 r32 b4 f2 c2a; do python bench.py --workload $w --also none --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | tail -1 >> gpurun_out/r2_bench_1gpu.jsonl; done
: > gpurun_out/r2_bench_1gpu_double_buffer.jsonl
for w in c2 c4 c5 x2 b4; do MB2_COMMIT_MODE=1 python bench.py --workload $w --also none --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | tail -1 >> gpurun_out/r2_bench_1gpu_double_buffer.jsonl; done
python bench.py --also none --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain_c2_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_c2_launches.csv python bench.py --also none --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2_launches.log 2>&1
# (the captured launch is the 13th of the fused kernel: generations late enough for the accepted fraction of the bench)
export MB2_COMMIT_MODE=2
for w in ${CAPTURE:-c2:de_step_best1bin_tma c5:de_step_best1bin_tma c4:de_step_rows_bulk x2:de_step_exp_tma}; do
  python bench.py --workload ${w%%:*} --also none --no-cpu-baseline --steps 2 --warmup 12 --sustain 0 > gpurun_out/plain_${w%%:*}_r2ip.log 2>&1 || { echo "plain run of ${w%%:*} failed"; continue; }
  ncu --set full --clock-control none --import-source on -k regex:${w#*:} --launch-skip 12 -c 1 -f -o gpurun_out/prof_${w%%:*}_r2ip python bench.py --workload ${w%%:*} --also none --no-cpu-baseline --steps 2 --warmup 12 --sustain 0 > gpurun_out/ncu_${w%%:*}_r2ip.log 2>&1
  ls -la gpurun_out/prof_${w%%:*}_r2ip.ncu-rep
done
unset MB2_COMMIT_MODE
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2.log 2>&1; tail -3 gpurun_out/smoke_r2.log
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_bench_default.json'))
print('default: c2 %.4f ms frac %.3f e2e %.4g' % (d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']), d['e2e']['breakdown'])
for k, v in d.get('also', {}).items():
    print(' ', k, v.get('ms_per_step'), v.get('roofline', {}).get('frac'), v.get('commit'), v.get('error'))
print(' cpu', {k: v for k, v in d.get('cpu_baseline', {}).items() if k != 'reference_python'})
for f in ('gpurun_out/r2_bench_1gpu.jsonl', 'gpurun_out/r2_bench_1gpu_double_buffer.jsonl'):
    print(f)
    for ln in open(f):
        try:
            e = json.loads(ln)
            print(e['config']['workload'][:40], '%.4f ms burst %.4f frac %.3f %s %s' % (e['ms_per_step'], e['details']['burst_ms_per_step'], e['roofline']['frac'], e['clocks']['sm_mhz'], e['roofline'].get('commit')))
        except Exception as ex:
            print('bad line', ex)
PY
