# round 2, call D: first hardware run of (1) the concurrent-search / lock-free-writer refactor of capi.cu, (2) the cp.async ring v3 scan
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
step() { echo "== $1"; }
step "stress v3";  timeout 300 python scripts/stress_v2.py 2>&1 | grep -E "stress|Error|error|assert" | tail -8; S3=${PIPESTATUS[0]}
echo "stress rc=$S3"
step "pytest -m gpu"
( timeout 1500 python -m pytest tests -x -q -m gpu --timeout 200 2>&1 | tail -40 ) > gpurun_out/pytest_gpu_d.log 2>&1; tail -12 gpurun_out/pytest_gpu_d.log
[ "$S3" != "0" ] && exit 1
VARS="GB200_SCAN_VARIANT=2;GB200_SCAN_THREADS=512;GB200_SCAN_THREADS=320;GB200_SCAN_THREADS=256;GB200_SCAN_CH=4;GB200_SCAN_CH=16;GB200_SCAN_HELP_MIN=2;GB200_SCAN_HELP_MIN=32;GB200_SCAN_ROWS=2"
step "headline + variants"
( timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --variants "$VARS" ) > gpurun_out/bench_headline_d.json 2> gpurun_out/bench_headline_d.err
grep -E "variant|recall|Error|error" gpurun_out/bench_headline_d.err | tail -16
python -c "import json; j=json.load(open('gpurun_out/bench_headline_d.json')); print('headline: QPS %.0f e2e %.0f ms/step %.4f scan kernel %.4f ms frac %.3f stages %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['roofline']['stage_ms']))"
step "ncu full, v3 scan"
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:ivfpq_scan -s 3 -c 1 -f -o gpurun_out/scan_prof_v3d \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > /dev/null 2> gpurun_out/ncu_full_d.err
tail -2 gpurun_out/ncu_full_d.err
