# gated run with tight timeouts: stress -> full GPU tests -> bench sweep (+ CTA timing)
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
timeout 240 python scripts/stress_v2.py 2>&1 | grep stress | tail -4; rc=${PIPESTATUS[0]}; echo "stress rc=$rc"; [ $rc -ne 0 ] && exit 1
if [ -n "$RUN_TESTS" ]; then
( timeout 400 python -m pytest tests -x -q -m gpu --timeout 150 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
grep -q passed gpurun_out/pytest_gpu.log || exit 1
fi
( timeout ${BENCH_TMO:-240} python bench.py --steps 10 --warmup 3 --no-cpu-baseline --variants "$VARIANTS" ) > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
grep -E "variant|recall" gpurun_out/bench_quick.err | tail -30
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_quick.json"))
print("QPS %.0f ms/step %.4f e2e %.0f recall %.4f" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["recall_at_10"]))
print("roofline", {k: j["roofline"][k] for k in ("achieved", "frac", "kernel_ms", "stage_ms")})
PY
if [ -n "$RUN_TIMING" ]; then
( GB200_SCAN_TIMING=1 timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --variants "$VARIANTS" ) > gpurun_out/bench_timing.json 2> gpurun_out/bench_timing.err
grep -E "variant|scan timing" gpurun_out/bench_timing.err | awk '/variant/{print last; print} {last=$0}' | tail -40
fi
