mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_ivfpq_gpu.py tests/test_golden_gpu.py -x -q -m gpu --timeout 180 2>&1 | tail -15 ) > gpurun_out/pytest_quick.log 2>&1
( GB200_SCAN_THREADS=384 timeout 300 python -m pytest tests/test_ivfpq_gpu.py -x -q -m gpu --timeout 180 2>&1 | tail -3 ) > gpurun_out/pytest_384.log 2>&1
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --variants "$VARIANTS" ) > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
tail -3 gpurun_out/pytest_quick.log; tail -2 gpurun_out/pytest_384.log; grep -E "variant|recall" gpurun_out/bench_quick.err; python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_quick.json"))
print("QPS %.0f ms/step %.4f e2e %.0f recall %.4f" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["recall_at_10"]))
print("roofline", {k: j["roofline"][k] for k in ("achieved", "frac", "kernel_ms", "stage_ms")})
PY
