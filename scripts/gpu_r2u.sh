# round 2, call AA: full GPU suite after the merged-batch fix and the fused exchange, concurrent searches
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
step() { echo "== $1"; }
step "pytest -m gpu"
( timeout 1500 python -m pytest tests -q -m gpu --timeout 400 2>&1 | grep -v "WARNING clustering" | tail -30 ) > gpurun_out/pytest_gpu_aa.log 2>&1; tail -5 gpurun_out/pytest_gpu_aa.log | cut -c1-300
step "plugin parity binary"
( timeout 600 gamma_b200/plugin/_build/plugin_parity ) > gpurun_out/plugin_parity_aa.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/plugin_parity_aa.log | cut -c1-200
step "concurrent search"
( timeout 600 python scripts/concurrent_bench.py ) > gpurun_out/concurrent_aa.json 2> gpurun_out/concurrent_aa.err
grep "^\[concurrent\]" gpurun_out/concurrent_aa.err | cut -c1-300; tail -1 gpurun_out/concurrent_aa.err | cut -c1-300; head -c 600 gpurun_out/concurrent_aa.json
