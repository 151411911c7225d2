# one gpurun call: GPU parity tests, smoke, headline bench.  Everything logs into gpurun_out/.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -x -q -m gpu -p pytest_timeout --timeout 180 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log 2>&1
( timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -5 ) > gpurun_out/smoke.log 2>&1
( timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 --cpu-queries ${CPUQ:-128} ${BENCH_ARGS} ) > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
