mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_flat_gpu.py -x -q -m gpu --timeout 150 2>&1 | tail -8 ) > gpurun_out/pytest_flat.log 2>&1; tail -3 gpurun_out/pytest_flat.log
( timeout 300 python scripts/bench_flat.py --steps 5 --warmup 2 ) > gpurun_out/bench_flat_c4.json 2> gpurun_out/bench_flat_c4.err
tail -2 gpurun_out/bench_flat_c4.err; cat gpurun_out/bench_flat_c4.json
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_gemm|flat_|tf32|row_norms" -c 60 --csv --log-file gpurun_out/launches_flat.csv \
    python scripts/bench_flat.py --N 1000000 --steps 1 --warmup 1 --check 4 ) > /dev/null 2> gpurun_out/ncu_launch_flat.err
python - <<'PY'
import csv
from collections import defaultdict
rows=list(csv.reader(open('gpurun_out/launches_flat.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
hdr=rows[hi]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value")
agg=defaultdict(list)
for r in rows[hi+1:]:
    if len(r)<=iv: continue
    try: agg[r[ik][:60]].append(float(r[iv].replace(',','')))
    except: pass
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    print("%-62s n=%3d mean=%9.1f us last=%9.1f us" % (k, len(v), sum(v)/len(v)/1e3, v[-1]/1e3))
PY
