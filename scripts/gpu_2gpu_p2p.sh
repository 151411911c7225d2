# round 2, call M (2 GPUs): exchange behind the C-ABI — 2-process test, headline at N=2 p2p
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
step() { echo "== $1"; }
step "2-GPU comm test"
( timeout 600 python -m pytest tests/test_comm_gpu.py -q -m gpu --timeout 500 2>&1 | tail -30 ) > gpurun_out/pytest_comm_m.log 2>&1; tail -8 gpurun_out/pytest_comm_m.log | cut -c1-600
step "headline N=2 exchange=p2p"
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --exchange p2p ) > gpurun_out/bench_headline_m2_p2p.json 2> gpurun_out/bench_headline_m2_p2p.err
grep -v "^\[W" gpurun_out/bench_headline_m2_p2p.err | tail -6 | cut -c1-300
python -c "
import json
l=[x for x in open('gpurun_out/bench_headline_m2_p2p.json').read().splitlines() if x.startswith('{')]
j=json.loads(l[-1]); print('N=2 p2p: QPS %.0f e2e %.0f ms/step %.4f' % (j['value'], j['e2e']['value'], j['ms_per_step']))"
