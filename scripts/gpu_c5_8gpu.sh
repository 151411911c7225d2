# round 2, call O (8 GPUs): config C5 (IVFPQ 100M nlist 65536 nprobe 64 batch 4096) query-sharded over 8 x B200
mkdir -p gpurun_out
( timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --workload c5 --steps 10 --warmup 3 ) > gpurun_out/bench_c5_n8.json 2> gpurun_out/bench_c5_n8.err
grep -v "^\[W" gpurun_out/bench_c5_n8.err | tail -8 | cut -c1-300
python -c "
import json
l=[x for x in open('gpurun_out/bench_c5_n8.json').read().splitlines() if x.startswith('{')]
j=json.loads(l[-1]); print('c5 N=8: QPS %.0f e2e %.0f ms/step %.3f scan %.3f frac %.3f recall %s' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['recall_at_10']))"
