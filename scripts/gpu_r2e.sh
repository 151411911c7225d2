# round 2, call E (re-run after the ring-epoch fix): TMA-fed posting ring in the v3 scan; coarse select from the GEMM's chunk minima
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
step() { echo "== $1"; }
step "stress v3 (TMA ring)";  timeout 300 python scripts/stress_v2.py 2>&1 | grep -E "stress|Error|error|assert" | tail -4; S3=${PIPESTATUS[0]}
echo "stress rc=$S3"
[ "$S3" != "0" ] && { echo "TMA ring failed the stress run: falling back to GB200_SCAN_TMA=0 for the rest"; export GB200_SCAN_TMA=0; }
step "pytest ivfpq + golden"
( timeout 1500 python -m pytest tests/test_ivfpq_gpu.py tests/test_golden_gpu.py -x -q -m gpu --timeout 200 2>&1 | tail -40 ) > gpurun_out/pytest_gpu_e.log 2>&1; tail -12 gpurun_out/pytest_gpu_e.log
VARS="GB200_SCAN_TMA=0;GB200_SCAN_TMA=1;GB200_SCAN_THREADS=320;GB200_SCAN_THREADS=256;GB200_SCAN_CH=4;GB200_SCAN_CH=16;GB200_COARSE_FULL_SELECT=1"
step "headline + variants"
( timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --variants "$VARS" ) > gpurun_out/bench_headline_e.json 2> gpurun_out/bench_headline_e.err
grep -E "variant|recall|Error|error" gpurun_out/bench_headline_e.err | tail -16
python -c "import json; j=json.load(open('gpurun_out/bench_headline_e.json')); print('headline: QPS %.0f e2e %.0f ms/step %.4f scan kernel %.4f ms frac %.3f stages %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['roofline']['stage_ms']))"
step "ncu launch list"
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ivfpq_scan|lut_build|probe_setup|rerank|coarse|tc_gemm|tf32|row_norms|build_valid" -c 150 --csv --log-file gpurun_out/launches_e.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > /dev/null 2> gpurun_out/ncu_launch_e.err
tail -2 gpurun_out/ncu_launch_e.err
step "ncu full, v3 scan"
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:ivfpq_scan -s 3 -c 1 -f -o gpurun_out/scan_prof_v3e \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > /dev/null 2> gpurun_out/ncu_full_e.err
tail -2 gpurun_out/ncu_full_e.err
