# round 2, call A: first hardware run of the v3 persistent scan (gated, every step under a timeout)
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
step() { echo "== $1"; }
step "stress v3";  timeout 300 python scripts/stress_v2.py 2>&1 | grep -E "stress|Error|error|assert" | tail -8; S3=${PIPESTATUS[0]}
echo "stress rc=$S3"
[ "$S3" = "124" ] && { echo "v3 HUNG in stress"; }
step "pytest -m gpu"
( timeout 1200 python -m pytest tests -x -q -m gpu --timeout 200 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
[ "$S3" != "0" ] && exit 1
VARS="GB200_SCAN_VARIANT=2;GB200_SCAN_THREADS=512;GB200_SCAN_THREADS=384;GB200_SCAN_CH=4;GB200_SCAN_CH=16;GB200_SCAN_HELP_MIN=2;GB200_SCAN_HELP_MIN=32;GB200_SCAN_ROWS=1;GB200_SCAN_ROWS=8;GB200_SCAN_PF=2;GB200_SCAN_PF=8;GB200_SCAN_PF=0"
step "headline + variants"
( timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --variants "$VARS" ) > gpurun_out/bench_headline.json 2> gpurun_out/bench_headline.err
grep -E "variant|recall" gpurun_out/bench_headline.err | tail -16
python -c "import json; j=json.load(open('gpurun_out/bench_headline.json')); print('headline: QPS %.0f e2e %.0f ms/step %.4f scan kernel %.4f ms frac %.3f stages %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['roofline']['stage_ms']))"
step "c3"
( timeout 400 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python -c "import json; j=json.load(open('gpurun_out/bench_c3.json')); print('c3: QPS %.0f ms/step %.4f scan kernel %.4f ms frac %.3f recall %s' % (j['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['recall_at_10']))"
step "c2"
( timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline --variants "GB200_SCAN_VARIANT=2;GB200_SCAN_ROWS=2;GB200_SCAN_ROWS=4;GB200_SCAN_CH=4" ) > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
grep -E "variant" gpurun_out/bench_c2.err | tail -5
python -c "import json; j=json.load(open('gpurun_out/bench_c2.json')); print('c2: QPS %.0f ms/step %.4f scan kernel %.4f ms frac %.3f stages %s' % (j['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['roofline']['stage_ms']))"
step "ncu launch list"
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ivfpq_scan|lut_build|probe_setup|rerank|coarse|tc_gemm|tf32|row_norms|build_valid" -c 150 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu_launch.err
tail -2 gpurun_out/ncu_launch.err
step "ncu full, v3 scan"
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:ivfpq_scan -s 3 -c 1 -f -o gpurun_out/scan_prof_v3 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > /dev/null 2> gpurun_out/ncu_full.err
tail -2 gpurun_out/ncu_full.err
ls -la gpurun_out | head -40
