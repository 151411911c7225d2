# v2 scan kernel: parity tests, then A/B of the tuning knobs on the headline workload
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
( timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
( timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --variants "$VARIANTS" ) > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
grep -E "variant|recall" gpurun_out/bench_quick.err | tail -30
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_quick.json"))
print("QPS %.0f ms/step %.4f e2e %.0f recall %.4f" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["recall_at_10"]))
print("roofline", {k: j["roofline"][k] for k in ("achieved", "frac", "kernel_ms", "stage_ms")})
PY
( GB200_SCAN_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_timing.json 2> gpurun_out/bench_timing.err
grep "scan timing" gpurun_out/bench_timing.err | tail -1
