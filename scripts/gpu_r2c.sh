# round 2, call C: v3 after removing the back-edge copy of the freshly loaded id
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
step() { echo "== $1"; }
step "stress v3";  timeout 300 python scripts/stress_v2.py 2>&1 | grep -E "stress|Error|error|assert" | tail -3; S3=${PIPESTATUS[0]}
echo "stress rc=$S3"
step "pytest subset"
( timeout 900 python -m pytest tests/test_ivfpq_gpu.py tests/test_golden_gpu.py -x -q -m gpu --timeout 200 -k "large_batch or full_search or filters or golden or m64 or no_rank" 2>&1 | tail -25 ) > gpurun_out/pytest_gpu_c.log 2>&1; tail -4 gpurun_out/pytest_gpu_c.log
[ "$S3" != "0" ] && exit 1
VARS="GB200_SCAN_VARIANT=2;GB200_SCAN_THREADS=512;GB200_SCAN_THREADS=384;GB200_SCAN_CH=4;GB200_SCAN_CH=6;GB200_SCAN_CH=12;GB200_SCAN_CH=16;GB200_SCAN_HELP_MIN=2;GB200_SCAN_HELP_MIN=32;GB200_SCAN_ROWS=1;GB200_SCAN_ROWS=2;GB200_SCAN_PF=0"
step "headline + variants"
( timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --variants "$VARS" ) > gpurun_out/bench_headline_c.json 2> gpurun_out/bench_headline_c.err
grep -E "variant|recall" gpurun_out/bench_headline_c.err | tail -16
python -c "import json; j=json.load(open('gpurun_out/bench_headline_c.json')); print('headline: QPS %.0f e2e %.0f ms/step %.4f scan kernel %.4f ms frac %.3f stages %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['roofline']['stage_ms']))"
step "ncu full, v3 scan"
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:ivfpq_scan -s 3 -c 1 -f -o gpurun_out/scan_prof_v3c \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > /dev/null 2> gpurun_out/ncu_full_c.err
tail -2 gpurun_out/ncu_full_c.err
