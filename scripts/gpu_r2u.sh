# round 2, call V: re-rank with raw rows staged in shared memory by cp.async (A/B against the load-on-use path): parity + headline
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
step() { echo "== $1"; }
step "pytest -m gpu"
( timeout 1800 python -m pytest tests -q -m gpu --timeout 300 2>&1 | grep -v "WARNING clustering" | tail -40 ) > gpurun_out/pytest_gpu_v.log 2>&1; tail -8 gpurun_out/pytest_gpu_v.log | cut -c1-300
step "headline + variants"
( timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --variants "GB200_RERANK_STAGE=0;GB200_RERANK_STAGE=1" ) > gpurun_out/bench_headline_v.json 2> gpurun_out/bench_headline_v.err
grep -E "variant|recall|Error|error" gpurun_out/bench_headline_v.err | tail -5
python -c "import json; j=json.load(open('gpurun_out/bench_headline_v.json')); print('headline: QPS %.0f e2e %.0f ms/step %.4f scan kernel %.4f ms frac %.3f stages %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['roofline']['stage_ms']))"
step "c5 quick (scale 0.02) for the wide-batch path"
( timeout 600 python bench.py --workload c5 --scale 0.02 --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_c5_v.json 2> gpurun_out/bench_c5_v.err
python -c "import json; j=json.load(open('gpurun_out/bench_c5_v.json')); print('c5 x0.02: QPS %.0f ms/step %.4f stages %s recall %s' % (j['value'], j['ms_per_step'], j['roofline']['stage_ms'], j['recall_at_10']))"
