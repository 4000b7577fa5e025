#!/bin/bash
# One-GPU evidence run of round 2 (under gpurun): the driver's bench command and the reference arm, the other
# workloads, the ncu launch list of the bench command, `ncu --set full` captures of the fused kernels, smoke().
# Everything lands in gpurun_out/; tools/summarize_ncu.py and tools/ncu_lines.py turn the reports into profiles/.
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
: > gpurun_out/r2_bench_1gpu.jsonl
for w in c2 c4 c5 c3 c3h x2 g128 r32 b4 f2 c2a; do python bench.py --workload $w --also none --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | tail -1 >> gpurun_out/r2_bench_1gpu.jsonl; done
python bench.py --also none --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain_c2_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_c2_launches.csv python bench.py --also none --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2_launches.log 2>&1
CAPTURE="${CAPTURE:-c2:de_step_best1bin_tma c5:de_step_best1bin_tma c4:de_step_rows_bulk c3:de_step_cheb_mma x2:de_step_exp_tma}" TAG=r2f tools/run_r2_profiles.sh
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2.log 2>&1; tail -3 gpurun_out/smoke_r2.log
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_bench_default.json'))
print('default: c2 %.4f ms frac %.3f e2e %.4g' % (d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']), d['e2e']['breakdown'])
for k, v in d.get('also', {}).items():
    print(' ', k, v.get('ms_per_step'), v.get('roofline', {}).get('frac'), v.get('error'))
print(' cpu', {k: v for k, v in d.get('cpu_baseline', {}).items() if k != 'reference_python'})
for ln in open('gpurun_out/r2_bench_1gpu.jsonl'):
    try:
        e = json.loads(ln)
        print(e['config']['workload'][:40], '%.4f ms burst %.4f frac %.3f %s' % (e['ms_per_step'], e['details']['burst_ms_per_step'], e['roofline']['frac'], e['clocks']['sm_mhz']))
    except Exception as ex:
        print('bad line', ex)
PY
