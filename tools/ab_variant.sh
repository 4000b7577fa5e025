for rep in 1 2; do
for lib in "" mystic_b200/_lib/libmb2_allwait.so; do
for wl in c2 x2 c5; do
R=$(MB2_LIBRARY=$lib timeout 120 python bench.py --workload $wl --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['config']['burst_ms_per_step'], d['clocks']['sm_mhz'])")
echo "lib=${lib:-default(leader wait)} $wl -> $R"
done; done; done
