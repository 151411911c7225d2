# round 2, call I: item order (full items first), table wait after the first blocks are requested; full suite
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
step() { echo "== $1"; }
step "stress v3";  timeout 300 python scripts/stress_v2.py 2>&1 | grep -E "stress|Error|error|assert" | tail -3; S3=${PIPESTATUS[0]}
echo "stress rc=$S3"
step "pytest -m gpu"
( timeout 1800 python -m pytest tests -q -m gpu --timeout 300 2>&1 | grep -v "WARNING clustering" | tail -80 ) > gpurun_out/pytest_gpu_i.log 2>&1; tail -15 gpurun_out/pytest_gpu_i.log
[ "$S3" != "0" ] && exit 1
VARS="GB200_SCAN_CH=4;GB200_SCAN_CH=6;GB200_SCAN_CH=12;GB200_SCAN_CH=16"
step "headline + variants"
( timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --variants "$VARS" ) > gpurun_out/bench_headline_i.json 2> gpurun_out/bench_headline_i.err
grep -E "variant|recall|Error|error" gpurun_out/bench_headline_i.err | tail -16
python -c "import json; j=json.load(open('gpurun_out/bench_headline_i.json')); print('headline: QPS %.0f e2e %.0f ms/step %.4f scan kernel %.4f ms frac %.3f stages %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['roofline']['stage_ms']))"
step "c3"
( timeout 400 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_c3_i.json 2> gpurun_out/bench_c3_i.err
python -c "import json; j=json.load(open('gpurun_out/bench_c3_i.json')); print('c3: QPS %.0f ms/step %.4f scan kernel %.4f ms frac %.3f recall %s' % (j['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['recall_at_10']))"
step "c2"
( timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_c2_i.json 2> gpurun_out/bench_c2_i.err
python -c "import json; j=json.load(open('gpurun_out/bench_c2_i.json')); print('c2: QPS %.0f ms/step %.4f scan kernel %.4f ms frac %.3f stages %s' % (j['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['roofline']['stage_ms']))"
