# round 2, call Y: full GPU suite (dynamic-smem attribute helper touches every launcher), headline, merged concurrent searches
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
step() { echo "== $1"; }
step "pytest -m gpu"
( timeout 1500 python -m pytest tests -q -m gpu --timeout 400 2>&1 | grep -v "WARNING clustering" | tail -30 ) > gpurun_out/pytest_gpu_y.log 2>&1; tail -5 gpurun_out/pytest_gpu_y.log | cut -c1-300
step "headline"
( timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/bench_headline_y.json 2> gpurun_out/bench_headline_y.err
python -c "import json; j=json.load(open('gpurun_out/bench_headline_y.json')); print('headline: QPS %.0f e2e %.0f ms/step %.4f scan kernel %.4f ms frac %.3f launches %s stages %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['gpu_launches'], j['roofline']['stage_ms']))"
step "concurrent search"
( timeout 600 python scripts/concurrent_bench.py ) > gpurun_out/concurrent_y.json 2> gpurun_out/concurrent_y.err
grep "^\[concurrent\]" gpurun_out/concurrent_y.err | cut -c1-300; tail -2 gpurun_out/concurrent_y.err | cut -c1-300; cat gpurun_out/concurrent_y.json
