# First GPU call of the next round: validate the two opt-in kernels written at the end of round 1 (no GPU time was left
# to run them), then A/B them.  Every step is gated and under a timeout; a hang costs at most that timeout.
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
step() { echo "== $1"; }
step "stress, default path";      timeout 240 python scripts/stress_v2.py 2>&1 | grep stress | tail -2; [ ${PIPESTATUS[0]} -ne 0 ] && exit 1
step "stress, work stealing";     GB200_SCAN_STEAL=1 timeout 240 python scripts/stress_v2.py 2>&1 | grep stress | tail -2; STEAL_OK=${PIPESTATUS[0]}
step "pytest, default suite";     ( timeout 500 python -m pytest tests -x -q -m gpu --timeout 150 2>&1 | tail -5 ) | tee gpurun_out/pytest_gpu.log | tail -2
if [ "$STEAL_OK" = "0" ]; then
  step "pytest, work stealing";   GB200_TEST_STEAL=1 timeout 300 python -m pytest tests/test_ivfpq_gpu.py -x -q -m gpu --timeout 120 -k stealing 2>&1 | tail -3
fi
step "pytest, M=64 kernel";       GB200_TEST_M64=1 timeout 300 python -m pytest tests/test_ivfpq_gpu.py -x -q -m gpu --timeout 120 -k m64 2>&1 | tail -3; M64_OK=${PIPESTATUS[0]}
VARS="GB200_COARSE_TIGHT=1;GB200_SCAN_THREADS=512"
[ "$STEAL_OK" = "0" ] && VARS="GB200_SCAN_STEAL=1;GB200_SCAN_STEAL=1,GB200_SCAN_TAIL=2;$VARS"
step "headline A/B: $VARS"
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --variants "$VARS" ) > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
grep -E "variant|recall" gpurun_out/bench_ab.err | tail -8
python -c "import json; j=json.load(open('gpurun_out/bench_ab.json')); print('default: QPS %.0f scan kernel %.4f ms frac %.3f' % (j['value'], j['roofline']['kernel_ms'], j['roofline']['frac']))"
step "C3 (PQ64 + filter): generic kernel vs M=64 kernel"
( timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_c3_generic.json 2> gpurun_out/bench_c3_generic.err
python -c "import json; j=json.load(open('gpurun_out/bench_c3_generic.json')); print('generic: QPS %.0f scan kernel %.4f ms frac %.3f' % (j['value'], j['roofline']['kernel_ms'], j['roofline']['frac']))"
if [ "$M64_OK" = "0" ]; then
  ( GB200_SCAN_M64=1 timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_c3_m64.json 2> gpurun_out/bench_c3_m64.err
  python -c "import json; j=json.load(open('gpurun_out/bench_c3_m64.json')); print('m64:     QPS %.0f scan kernel %.4f ms frac %.3f recall %.4f' % (j['value'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['recall_at_10']))"
fi
