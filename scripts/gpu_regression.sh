# final regression of the round: plugin binary, smoke, all GPU tests, headline bench with the CPU baseline, reference arm,
# ncu launch list, workloads c2 / c3
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
T=${TAG:-final}
step() { echo "== $1"; }
step "plugin parity binary"
( timeout 600 gamma_b200/plugin/_build/plugin_parity ) > gpurun_out/plugin_parity_$T.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/plugin_parity_$T.log | cut -c1-200
step "smoke"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
step "pytest -m gpu"
( timeout 1800 python -m pytest tests -q -m gpu --timeout 300 2>&1 | grep -v "WARNING clustering" | tail -60 ) > gpurun_out/pytest_gpu_$T.log 2>&1; tail -3 gpurun_out/pytest_gpu_$T.log | cut -c1-300
step "headline (with cpu baseline)"
( timeout 900 python bench.py ) > gpurun_out/bench_headline_$T.json 2> gpurun_out/bench_headline_$T.err
grep -E "Error|error" gpurun_out/bench_headline_$T.err | tail -5
python -c "import json; j=json.load(open('gpurun_out/bench_headline_$T.json')); print('headline: QPS %.0f e2e %.0f ms/step %.4f scan kernel %.4f ms frac %.3f stages %s cpu %s launches %s clocks %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['roofline']['stage_ms'], j['cpu_baseline'], j['gpu_launches'], j['clocks']))"
step "reference arm"
( timeout 600 python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/bench_reference_$T.json 2> gpurun_out/bench_reference_$T.err
python -c "import json; j=json.load(open('gpurun_out/bench_reference_$T.json')); print('reference: QPS %.0f cores %s' % (j['value'], j['cpu_baseline']['cores']))"
step "ncu launch list"
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ivfpq_scan|lut_build|probe_setup|rerank|coarse|tc_gemm|tf32|row_norms|rows_prep|build_valid|linear" -c 150 --csv --log-file gpurun_out/launches_$T.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > /dev/null 2> gpurun_out/ncu_launch_$T.err
tail -1 gpurun_out/ncu_launch_$T.err
for wl in c2 c3; do
step "workload $wl"
( timeout 600 python bench.py --workload $wl --no-cpu-baseline ) > gpurun_out/bench_${wl}_$T.json 2> gpurun_out/bench_${wl}_$T.err
python -c "import json; j=json.load(open('gpurun_out/bench_${wl}_$T.json')); print('$wl: QPS %.0f e2e %.0f ms/step %.4f scan kernel %.4f ms frac %.3f recall %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['recall_at_10']))"
done
