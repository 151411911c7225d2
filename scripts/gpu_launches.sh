mkdir -p gpurun_out
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"coarse|tc_gemm|tf32|row_norms|lut_build|ivfpq_scan|rerank|query_order|build_valid" -c 200 --csv --log-file gpurun_out/launches_mine.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_under_ncu4.json 2> gpurun_out/ncu_launch4.err
python - <<'PY'
import csv
from collections import defaultdict
rows=list(csv.reader(open('gpurun_out/launches_mine.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
hdr=rows[hi]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value")
agg=defaultdict(list)
for r in rows[hi+1:]:
    if len(r)<=iv: continue
    try: agg[r[ik][:60]].append(float(r[iv].replace(',','')))
    except: pass
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    print("%-62s n=%3d mean=%8.1f us" % (k, len(v), sum(v)/len(v)/1e3))
PY
