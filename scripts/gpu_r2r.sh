# round 2, call R: scan with next-query prefetch + early table request; candidate-buffer size sweep; comm test on one GPU
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
step() { echo "== $1"; }
step "stress v3";  timeout 300 python scripts/stress_v2.py 2>&1 | grep -E "stress|Error|error|assert" | tail -3; S3=${PIPESTATUS[0]}
echo "stress rc=$S3"
step "pytest ivfpq + golden + comm"
( timeout 1500 python -m pytest tests/test_ivfpq_gpu.py tests/test_golden_gpu.py tests/test_comm_gpu.py -q -m gpu --timeout 300 2>&1 | grep -v "WARNING clustering" | tail -40 ) > gpurun_out/pytest_gpu_r.log 2>&1; tail -12 gpurun_out/pytest_gpu_r.log | cut -c1-400
[ "$S3" != "0" ] && exit 1
VARS="GB200_SCAN_CAP=1536;GB200_SCAN_CAP=1280;GB200_SCAN_CAP=1536,GB200_SCAN_CH=6;GB200_SCAN_HELP_MIN=16"
step "headline + variants"
( timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --variants "$VARS" ) > gpurun_out/bench_headline_r.json 2> gpurun_out/bench_headline_r.err
grep -E "variant|recall|Error|error" gpurun_out/bench_headline_r.err | tail -16
python -c "import json; j=json.load(open('gpurun_out/bench_headline_r.json')); print('headline: QPS %.0f e2e %.0f ms/step %.4f scan kernel %.4f ms frac %.3f stages %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['roofline']['stage_ms']))"
step "ncu full, scan"
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ivfpq_scan" -s 3 -c 1 -f -o gpurun_out/scan_prof_r \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > /dev/null 2> gpurun_out/ncu_full_r.err
tail -2 gpurun_out/ncu_full_r.err
