# round 2 (2 GPUs): exchange fused into the re-rank kernel — 2-process test on two devices, headline at N=2 (deferred wait)
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
step() { echo "== $1"; }
step "2-GPU comm test"
( timeout 400 python -m pytest tests/test_comm_gpu.py -q -m gpu --timeout 300 2>&1 | tail -30 ) > gpurun_out/pytest_comm_n2.log 2>&1; tail -5 gpurun_out/pytest_comm_n2.log | cut -c1-600
for ex in p2p-deferred; do
step "headline N=2 exchange=$ex"
( timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --exchange $ex ) > gpurun_out/bench_headline_n2_$ex.json 2> gpurun_out/bench_headline_n2_$ex.err
grep -v "^\[W\|^W0\|^\*\*\*" gpurun_out/bench_headline_n2_$ex.err | tail -4 | cut -c1-300
python -c "
import json
l=[x for x in open('gpurun_out/bench_headline_n2_$ex.json').read().splitlines() if x.startswith('{')]
j=json.loads(l[-1]); print('N=2 $ex: QPS %.0f e2e %.0f ms/step %.4f launches %s stages %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['gpu_launches'], j['roofline']['stage_ms']))"
done
