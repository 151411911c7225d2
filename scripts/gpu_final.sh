# checkpoint run: GPU tests, headline bench (with CPU baseline), ncu launch list, ncu full capture of the scan kernel
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
( timeout 500 python -m pytest tests -x -q -m gpu --timeout 150 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
grep -q passed gpurun_out/pytest_gpu.log || exit 1
( timeout 500 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_headline.json 2> gpurun_out/bench_headline.err
tail -2 gpurun_out/bench_headline.err; cat gpurun_out/bench_headline.json
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ivfpq_scan|lut_build|probe_setup|plan_items|rerank|coarse|tc_gemm|tf32|row_norms|query_order|build_valid" -c 150 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu_launch.err
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:ivfpq_scan -s 3 -c 1 -f -o gpurun_out/scan_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > /dev/null 2> gpurun_out/ncu_full.err
tail -2 gpurun_out/ncu_launch.err; tail -2 gpurun_out/ncu_full.err; ls -la gpurun_out | head -30
