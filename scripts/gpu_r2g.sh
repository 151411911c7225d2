# round 2, call G: v3 loop with producer queue + item table, new GEMM epilogue, fused rows_prep; full GPU suite
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
step() { echo "== $1"; }
step "stress v3";  timeout 300 python scripts/stress_v2.py 2>&1 | grep -E "stress|Error|error|assert" | tail -3; S3=${PIPESTATUS[0]}
echo "stress rc=$S3"
step "pytest -m gpu"
( timeout 1800 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -80 ) > gpurun_out/pytest_gpu_g.log 2>&1; tail -15 gpurun_out/pytest_gpu_g.log
[ "$S3" != "0" ] && exit 1
VARS="GB200_SCAN_CH=4;GB200_SCAN_CH=6;GB200_SCAN_CH=12;GB200_SCAN_THREADS=320;GB200_SCAN_HELP_MIN=4"
step "headline + variants"
( timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --variants "$VARS" ) > gpurun_out/bench_headline_g.json 2> gpurun_out/bench_headline_g.err
grep -E "variant|recall|Error|error" gpurun_out/bench_headline_g.err | tail -16
python -c "import json; j=json.load(open('gpurun_out/bench_headline_g.json')); print('headline: QPS %.0f e2e %.0f ms/step %.4f scan kernel %.4f ms frac %.3f stages %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['roofline']['stage_ms']))"
step "ncu launch list"
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ivfpq_scan|lut_build|probe_setup|rerank|coarse|tc_gemm|tf32|row_norms|rows_prep|build_valid" -c 150 --csv --log-file gpurun_out/launches_g.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > /dev/null 2> gpurun_out/ncu_launch_g.err
tail -2 gpurun_out/ncu_launch_g.err
step "ncu full, scan + gemm"
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ivfpq_scan|tc_gemm" -s 6 -c 2 -f -o gpurun_out/prof_g \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > /dev/null 2> gpurun_out/ncu_full_g.err
tail -2 gpurun_out/ncu_full_g.err
