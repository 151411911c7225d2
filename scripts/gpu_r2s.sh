# round 2, call S: scan flags sweep (bit0 late fetch of the next query inside the scan, bit1 early tables, bit2 approximate in-loop prunes)
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
step() { echo "== $1"; }
step "stress, all flags on";  GB200_SCAN_FLAGS=7 timeout 300 python scripts/stress_v2.py 2>&1 | grep -E "stress|Error|error|assert" | tail -3; S3=${PIPESTATUS[0]}
echo "stress rc=$S3"
[ "$S3" != "0" ] && exit 1
step "pytest subset, flags 5"
( GB200_SCAN_FLAGS=5 timeout 900 python -m pytest tests/test_ivfpq_gpu.py tests/test_golden_gpu.py -x -q -m gpu --timeout 200 -k "large_batch or full_search or filters or golden or no_rank or many_lists" 2>&1 | tail -5 ) > gpurun_out/pytest_gpu_s.log 2>&1; tail -3 gpurun_out/pytest_gpu_s.log
VARS="GB200_SCAN_FLAGS=1;GB200_SCAN_FLAGS=4;GB200_SCAN_FLAGS=5;GB200_SCAN_FLAGS=3;GB200_SCAN_FLAGS=7;GB200_SCAN_FLAGS=5,GB200_SCAN_CAP=1536;GB200_SCAN_FLAGS=0"
step "headline + variants"
( timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --variants "$VARS" ) > gpurun_out/bench_headline_s.json 2> gpurun_out/bench_headline_s.err
grep -E "variant|recall|Error|error" gpurun_out/bench_headline_s.err | tail -16
python -c "import json; j=json.load(open('gpurun_out/bench_headline_s.json')); print('headline: QPS %.0f e2e %.0f ms/step %.4f scan kernel %.4f ms frac %.3f' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac']))"
