# round 2, call T: ncu of the re-rank kernel (headline) and of the scan on config C2 (batch 256 < CTA slots)
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:"rerank" -s 3 -c 1 -f -o gpurun_out/rerank_prof_t \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > /dev/null 2> gpurun_out/ncu_rerank_t.err
tail -2 gpurun_out/ncu_rerank_t.err
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ivfpq_scan" -s 3 -c 1 -f -o gpurun_out/scan_c2_prof_t \
    python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline ) > /dev/null 2> gpurun_out/ncu_c2_t.err
tail -2 gpurun_out/ncu_c2_t.err
( timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline --variants "GB200_SCAN_ROWS=8;GB200_SCAN_ROWS=8,GB200_SCAN_HELP_MIN=2;GB200_SCAN_ROWS=4,GB200_SCAN_CH=4,GB200_SCAN_HELP_MIN=2" ) > gpurun_out/bench_c2_t.json 2> gpurun_out/bench_c2_t.err
grep -E "variant" gpurun_out/bench_c2_t.err | tail -5
python -c "import json; j=json.load(open('gpurun_out/bench_c2_t.json')); print('c2: QPS %.0f ms/step %.4f scan kernel %.4f ms frac %.3f stages %s' % (j['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['roofline']['stage_ms']))"
