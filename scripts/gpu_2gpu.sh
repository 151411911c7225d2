# round 2, call L (2 GPUs): the exchange behind the C-ABI — 2-process test, headline at N=2 with p2p and with NCCL, c4 at N=2
mkdir -p gpurun_out
export GB200_BENCH_CACHE=/tmp/gbcache
step() { echo "== $1"; }
step "2-GPU comm test"
( timeout 600 python -m pytest tests/test_comm_gpu.py -q -m gpu --timeout 500 2>&1 | tail -30 ) > gpurun_out/pytest_comm_l.log 2>&1; tail -8 gpurun_out/pytest_comm_l.log
step "headline N=1 (fills the cache)"
( timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/bench_headline_l1.json 2> gpurun_out/bench_headline_l1.err
python -c "import json; j=json.load(open('gpurun_out/bench_headline_l1.json')); print('N=1: QPS %.0f e2e %.0f ms/step %.4f' % (j['value'], j['e2e']['value'], j['ms_per_step']))"
for EX in p2p nccl; do
  step "headline N=2 exchange=$EX"
  ( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --exchange $EX ) > gpurun_out/bench_headline_l2_$EX.json 2> gpurun_out/bench_headline_l2_$EX.err
  tail -3 gpurun_out/bench_headline_l2_$EX.err | cut -c1-300
  python -c "import json; j=json.load(open('gpurun_out/bench_headline_l2_$EX.json')); print('N=2 $EX: QPS %.0f e2e %.0f ms/step %.4f' % (j['value'], j['e2e']['value'], j['ms_per_step']))"
done
step "c4 N=2"
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c4 --steps 5 --warmup 3 ) > gpurun_out/bench_c4_l2.json 2> gpurun_out/bench_c4_l2.err
tail -3 gpurun_out/bench_c4_l2.err | cut -c1-300
python -c "import json; j=json.load(open('gpurun_out/bench_c4_l2.json')); print('c4 N=2: QPS %.0f e2e %.0f ms/step %.3f' % (j['value'], j['e2e']['value'], j['ms_per_step']))"
