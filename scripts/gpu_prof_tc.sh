mkdir -p gpurun_out
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm|coarse_select_row|rerank_kernel" -s 9 -c 3 -f -o gpurun_out/tc_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > /dev/null 2> gpurun_out/ncu_tc.err
tail -3 gpurun_out/ncu_tc.err
