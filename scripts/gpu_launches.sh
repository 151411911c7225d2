# ncu launch list (my kernels only) for each env setting in $SETS (';'-separated, ','-joined ENV=VAL)
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
i=0
IFS=';' read -ra ARR <<< "$SETS"
for s in "${ARR[@]}"; do
  i=$((i+1))
  envs=$(echo "$s" | tr ',' ' ')
  ( env $envs timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ivfpq_scan|lut_build|probe_setup|plan_items|rerank|coarse|tc_gemm|tf32|row_norms|query_order|build_valid" -c 120 --csv --log-file gpurun_out/launches_$i.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > /dev/null 2> gpurun_out/ncu_launch_$i.err
  echo "== set $i: $s"
  python - <<PY
import csv
from collections import defaultdict
rows=list(csv.reader(open('gpurun_out/launches_$i.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
hdr=rows[hi]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value")
agg=defaultdict(list)
for r in rows[hi+1:]:
    if len(r)<=iv: continue
    try: agg[r[ik][:60]].append(float(r[iv].replace(',','')))
    except: pass
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    print("%-62s n=%3d last=%8.1f us" % (k, len(v), v[-1]/1e3))
PY
done
