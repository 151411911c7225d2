# GPU tests + ncu captures of the headline bench command (one GPU).  Logs into gpurun_out/.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu --timeout 180 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log 2>&1
# launch list (per-launch device time; shares, not absolutes)
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu_launch.err
# full capture of the dominant kernel
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:ivfpq_scan -s 3 -c 2 -f -o gpurun_out/scan_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > /dev/null 2> gpurun_out/ncu_full.err
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/ncu_launch.err; tail -3 gpurun_out/ncu_full.err; ls -la gpurun_out
