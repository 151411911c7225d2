# round 2, call N (4 GPUs): config C4 (FLAT IP d=768 5M batch 512) query-sharded over 4 x B200
mkdir -p gpurun_out
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --workload c4 --steps 5 --warmup 3 ) > gpurun_out/bench_c4_n4.json 2> gpurun_out/bench_c4_n4.err
grep -v "^\[W" gpurun_out/bench_c4_n4.err | tail -4 | cut -c1-300
python -c "
import json
l=[x for x in open('gpurun_out/bench_c4_n4.json').read().splitlines() if x.startswith('{')]
j=json.loads(l[-1]); print('c4 N=4: QPS %.0f e2e %.0f ms/step %.3f' % (j['value'], j['e2e']['value'], j['ms_per_step']))"
