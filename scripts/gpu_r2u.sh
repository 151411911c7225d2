# round 2, last call: the concurrency tests and a short headline run on the build with the header-only batching policy
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
step() { echo "== $1"; }
step "pytest realtime + comm"
( timeout 400 python -m pytest tests/test_realtime_gpu.py tests/test_comm_gpu.py -q -m gpu --timeout 300 2>&1 | grep -v "WARNING clustering" | tail -30 ) > gpurun_out/pytest_gpu_ab.log 2>&1; tail -3 gpurun_out/pytest_gpu_ab.log | cut -c1-300
step "headline"
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_headline_ab.json 2> gpurun_out/bench_headline_ab.err
python -c "import json; j=json.load(open('gpurun_out/bench_headline_ab.json')); print('headline: QPS %.0f e2e %.0f ms/step %.4f scan kernel %.4f ms frac %.3f' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac']))"
step "plugin parity binary"
( timeout 200 gamma_b200/plugin/_build/plugin_parity ) > gpurun_out/plugin_parity_ab.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/plugin_parity_ab.log | cut -c1-200
