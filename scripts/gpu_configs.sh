# tests (incl. large-batch plan parity), FLAT C4 bench + ncu evidence, C3 / C2 bench lines
mkdir -p gpurun_out
( timeout 500 python -m pytest tests -x -q -m gpu --timeout 150 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
( timeout 300 python scripts/bench_flat.py --steps 5 --warmup 2 ) > gpurun_out/bench_flat_c4.json 2> gpurun_out/bench_flat_c4.err
tail -3 gpurun_out/bench_flat_c4.err; cat gpurun_out/bench_flat_c4.json
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_gemm|flat_|tf32|row_norms" -c 60 --csv --log-file gpurun_out/launches_flat.csv \
    python scripts/bench_flat.py --N 1000000 --steps 1 --warmup 1 --check 4 ) > /dev/null 2> gpurun_out/ncu_launch_flat.err
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 8 -c 1 -f -o gpurun_out/flat_gemm_prof \
    python scripts/bench_flat.py --N 1000000 --steps 1 --warmup 1 --check 4 ) > /dev/null 2> gpurun_out/ncu_full_flat.err
tail -2 gpurun_out/ncu_full_flat.err
export GB200_BENCH_CACHE=/tmp/gbcache
( timeout 400 python bench.py --workload c3 --steps 10 --warmup 3 --cpu-queries 128 ) > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
tail -2 gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json
( timeout 300 python bench.py --workload c2 --steps 20 --warmup 5 ) > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
tail -2 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
ls -la gpurun_out | head -30
