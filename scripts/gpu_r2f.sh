# round 2, call F: full GPU suite (realtime / plugin additions), 416- and 448-thread scan shapes
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
step() { echo "== $1"; }
step "plugin parity binary"
( timeout 600 gamma_b200/plugin/_build/plugin_parity ) > gpurun_out/plugin_parity_f.log 2>&1; echo "rc=$?"; cat gpurun_out/plugin_parity_f.log | tail -25
step "pytest -m gpu"
( timeout 1800 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -60 ) > gpurun_out/pytest_gpu_f.log 2>&1; tail -30 gpurun_out/pytest_gpu_f.log
VARS="GB200_SCAN_THREADS=448;GB200_SCAN_THREADS=416;GB200_SCAN_THREADS=448,GB200_SCAN_CH=6;GB200_SCAN_TMA=1"
step "headline + variants"
( timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --variants "$VARS" ) > gpurun_out/bench_headline_f.json 2> gpurun_out/bench_headline_f.err
grep -E "variant|recall|Error|error" gpurun_out/bench_headline_f.err | tail -16
python -c "import json; j=json.load(open('gpurun_out/bench_headline_f.json')); print('headline: QPS %.0f e2e %.0f ms/step %.4f scan kernel %.4f ms frac %.3f stages %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['roofline']['stage_ms']))"
step "ncu full, coarse GEMM + select"
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm|coarse_select" -s 6 -c 2 -f -o gpurun_out/coarse_prof_f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > /dev/null 2> gpurun_out/ncu_coarse_f.err
tail -2 gpurun_out/ncu_coarse_f.err
