#!/bin/bash
# One-GPU evidence run after K2p (under gpurun): the whole GPU suite, the driver's bench command, per-workload lines,
# the ncu launch list of the bench command and an `ncu --set full` capture of K2p at BASELINE config 2, smoke().
# Everything lands in gpurun_out/; tools/summarize_ncu.py and tools/ncu_lines.py turn the reports into profiles/.
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c_gpu_tests.log 2>&1; tail -3 gpurun_out/r2c_gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err
: > gpurun_out/r2c_bench_1gpu.jsonl
for w in c2 c5 x2 o2; do timeout 120 python bench.py --workload $w --also none --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | tail -1 >> gpurun_out/r2c_bench_1gpu.jsonl; done
timeout 120 python bench.py --also none --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain_c2_launches_r2c.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_c2_launches.csv python bench.py --also none --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2_launches_r2c.log 2>&1
timeout 120 python bench.py --workload c2 --also none --no-cpu-baseline --steps 2 --warmup 12 --sustain 0 > gpurun_out/plain_c2_r2c.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:de_step_best1bin_pool --launch-skip 12 -c 1 -f -o gpurun_out/prof_c2_r2c python bench.py --workload c2 --also none --no-cpu-baseline --steps 2 --warmup 12 --sustain 0 > gpurun_out/ncu_c2_r2c.log 2>&1
ls -la gpurun_out/prof_c2_r2c.ncu-rep
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2c.log 2>&1; tail -3 gpurun_out/smoke_r2c.log
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2c_bench_default.json'))
print('default: c2 %.4f ms frac %.3f e2e %.4g' % (d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']), d['e2e']['breakdown'])
print(' ', d['roofline'].get('kernel'), d['roofline'].get('traffic_source'))
for k, v in d.get('also', {}).items():
    print(' ', k, v.get('ms_per_step'), v.get('roofline', {}).get('frac'), v.get('commit'), v.get('error'))
print(' cpu', {k: v for k, v in d.get('cpu_baseline', {}).items() if k != 'reference_python'})
for ln in open('gpurun_out/r2c_bench_1gpu.jsonl'):
    try:
        e = json.loads(ln)
        print(e['config']['workload'][:40], '%.4f ms burst %.4f frac %.3f burst_frac %.3f %s %s' % (e['ms_per_step'], e['details']['burst_ms_per_step'], e['roofline']['frac'], e['roofline'].get('frac_burst', 0), e['clocks']['sm_mhz'], e['roofline'].get('kernel')))
    except Exception as ex:
        print('bad line', ex)
PY
