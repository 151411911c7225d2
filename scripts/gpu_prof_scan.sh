# ncu --set full capture of the scan kernel for each env setting in $SETS (';'-separated, ','-joined ENV=VAL)
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
i=0
IFS=';' read -ra ARR <<< "$SETS"
for s in "${ARR[@]}"; do
  i=$((i+1))
  envs=$(echo "$s" | tr ',' ' ')
  ( env $envs timeout 600 ncu --set full --clock-control none --import-source on -k regex:ivfpq_scan -s 3 -c 1 -f -o gpurun_out/scan_prof_$i \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > /dev/null 2> gpurun_out/ncu_full_$i.err
  echo "set $i: $s"; tail -2 gpurun_out/ncu_full_$i.err
done
ls -la gpurun_out
