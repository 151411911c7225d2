mkdir -p gpurun_out
( timeout 400 python scripts/debug_large_r.py ) > gpurun_out/debug_large_r.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:ivfpq_scan_m32 -s 3 -c 1 -f -o gpurun_out/scan_prof3 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > /dev/null 2> gpurun_out/ncu_full3.err
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_under_ncu3.json 2> gpurun_out/ncu_launch3.err
grep -v "^WARNING" gpurun_out/debug_large_r.log | tail -30
