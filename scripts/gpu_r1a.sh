# round-1 re-entry: GPU tests, headline bench (with CPU baseline), ncu launch list, ncu full capture of the scan kernel
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
( timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
( timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_headline.json 2> gpurun_out/bench_headline.err
tail -3 gpurun_out/bench_headline.err; cat gpurun_out/bench_headline.json
( GB200_SCAN_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_timing.json 2> gpurun_out/bench_timing.err
grep "scan timing" gpurun_out/bench_timing.err | tail -2
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu_launch.err
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:ivfpq_scan -s 3 -c 1 -f -o gpurun_out/scan_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > /dev/null 2> gpurun_out/ncu_full.err
tail -3 gpurun_out/ncu_launch.err; tail -3 gpurun_out/ncu_full.err; ls -la gpurun_out
