# round 2, call Z: full regression (plugin binary + all GPU tests + smoke), headline bench with the CPU baseline, ncu captures
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
step() { echo "== $1"; }
step "plugin parity binary"
( timeout 600 gamma_b200/plugin/_build/plugin_parity ) > gpurun_out/plugin_parity_z.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/plugin_parity_z.log | cut -c1-200
step "smoke"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
step "pytest -m gpu"
( timeout 1800 python -m pytest tests -q -m gpu --timeout 300 2>&1 | grep -v "WARNING clustering" | tail -60 ) > gpurun_out/pytest_gpu_z.log 2>&1; tail -12 gpurun_out/pytest_gpu_z.log
step "headline (with cpu baseline)"
( timeout 900 python bench.py ) > gpurun_out/bench_headline_z.json 2> gpurun_out/bench_headline_z.err
grep -E "recall|Error|error" gpurun_out/bench_headline_z.err | tail -5
python -c "import json; j=json.load(open('gpurun_out/bench_headline_z.json')); print('headline: QPS %.0f e2e %.0f ms/step %.4f scan kernel %.4f ms frac %.3f stages %s cpu %s launches %s clocks %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['roofline']['stage_ms'], j['cpu_baseline'], j['gpu_launches'], j['clocks']))"
step "reference arm"
( timeout 600 python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/bench_reference_z.json 2> gpurun_out/bench_reference_z.err
python -c "import json; j=json.load(open('gpurun_out/bench_reference_z.json')); print('reference: QPS %.0f cores %s' % (j['value'], j['cpu_baseline']['cores']))"
step "ncu launch list"
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ivfpq_scan|lut_build|probe_setup|rerank|coarse|tc_gemm|tf32|row_norms|rows_prep|build_valid|linear" -c 150 --csv --log-file gpurun_out/launches_z.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > /dev/null 2> gpurun_out/ncu_launch_z.err
tail -2 gpurun_out/ncu_launch_z.err
step "ncu full, scan"
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ivfpq_scan" -s 3 -c 1 -f -o gpurun_out/scan_prof_z \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > /dev/null 2> gpurun_out/ncu_full_z.err
tail -2 gpurun_out/ncu_full_z.err
step "concurrent search"
( timeout 600 python scripts/concurrent_bench.py ) > gpurun_out/concurrent_z.json 2> gpurun_out/concurrent_z.err
grep "^\[concurrent\]" gpurun_out/concurrent_z.err | cut -c1-300; tail -1 gpurun_out/concurrent_z.err | cut -c1-300; head -c 700 gpurun_out/concurrent_z.json
