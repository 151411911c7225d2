mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_ivfpq_gpu.py -x -q -m gpu -k "coarse" --timeout 100 2>&1 | tail -25 ) > gpurun_out/pytest_coarse_tc.log 2>&1
tail -25 gpurun_out/pytest_coarse_tc.log
if grep -q "passed" gpurun_out/pytest_coarse_tc.log && ! grep -q "failed" gpurun_out/pytest_coarse_tc.log; then
  ( timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log 2>&1
  tail -4 gpurun_out/pytest_gpu.log
  ( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --variants "GB200_COARSE=simt;GB200_SCAN_SPLITS=2" ) > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
  grep -E "variant|recall" gpurun_out/bench_quick.err | tail -6
  python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_quick.json"))
print("QPS %.0f ms/step %.4f e2e %.0f recall %.4f" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["recall_at_10"]))
print("roofline", {k: j["roofline"][k] for k in ("achieved", "frac", "kernel_ms", "stage_ms")})
PY
else
  echo "coarse tc test failed; rerun with simt to confirm the rest"
  ( GB200_COARSE=simt timeout 300 python -m pytest tests/test_ivfpq_gpu.py tests/test_plugin_gpu.py -x -q -m gpu --timeout 200 2>&1 | tail -8 )
fi
