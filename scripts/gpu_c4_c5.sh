# round 2, call J: bench workloads c4 (flat) and c5 (100M, device-built): small-scale dry runs, then the full shapes
mkdir -p gpurun_out
step() { echo "== $1"; }
step "c4 dry run (scale 0.04)"
( timeout 600 python bench.py --workload c4 --scale 0.04 --steps 3 --warmup 3 ) > gpurun_out/bench_c4_dry.json 2> gpurun_out/bench_c4_dry.err; RC4=$?
tail -5 gpurun_out/bench_c4_dry.err | cut -c1-300; head -c 600 gpurun_out/bench_c4_dry.json; echo " rc=$RC4"
step "c5 dry run (scale 0.01)"
( timeout 900 python bench.py --workload c5 --scale 0.01 --steps 3 --warmup 3 --cpu-queries 32 ) > gpurun_out/bench_c5_dry.json 2> gpurun_out/bench_c5_dry.err; RC5=$?
tail -8 gpurun_out/bench_c5_dry.err | cut -c1-300; head -c 600 gpurun_out/bench_c5_dry.json; echo " rc=$RC5"
if [ "$RC4" = "0" ]; then
  step "c4 full"
  ( timeout 900 python bench.py --workload c4 --steps 5 --warmup 3 ) > gpurun_out/bench_c4_j.json 2> gpurun_out/bench_c4_j.err
  tail -4 gpurun_out/bench_c4_j.err | cut -c1-300
  python -c "import json; j=json.load(open('gpurun_out/bench_c4_j.json')); print('c4: QPS %.0f e2e %.0f ms/step %.3f roofline %.3f (alg %.0f TFLOP/s) cpu %s check %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['frac'], j['roofline']['achieved'], j['cpu_baseline'], j['check']))"
fi
if [ "$RC5" = "0" ]; then
  step "c5 full, 1 GPU"
  ( timeout 1500 python bench.py --workload c5 --steps 10 --warmup 3 --cpu-queries 64 ) > gpurun_out/bench_c5_j.json 2> gpurun_out/bench_c5_j.err
  tail -8 gpurun_out/bench_c5_j.err | cut -c1-300
  python -c "import json; j=json.load(open('gpurun_out/bench_c5_j.json')); print('c5: QPS %.0f e2e %.0f ms/step %.3f scan %.4f ms frac %.3f recall %s stages %s cpu %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['recall_at_10'], j['roofline']['stage_ms'], j['cpu_baseline']))"
fi
