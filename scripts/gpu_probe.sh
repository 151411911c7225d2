mkdir -p gpurun_out
nproc > gpurun_out/probe.log
python - >> gpurun_out/probe.log 2>&1 <<'PY'
import faulthandler, sys, time, json
faulthandler.dump_traceback_later(150, exit=True)
sys.path.insert(0, ".")
import numpy as np
t=time.time()
from gamma_b200 import api
print("devices", api.lib().gb200_device_count(), time.time()-t, flush=True)
ix = api.B200FLAT(0)
print("flat init", ix.Init('{"metric_type":"L2"}', 16), flush=True)
ix.Add(np.eye(3,16,dtype=np.float32))
print(ix.Search(np.zeros((2,16),np.float32), 5, metric="L2"), flush=True)
from oracle import ref
print("threads", ref.max_threads(), flush=True)
sys.path.insert(0, "tests")
from conftest import get_ref_fixture
t=time.time()
f = get_ref_fixture("l2_m16", N=20000, d=64, nlist=64, M=16, metric="L2", nq=48, n_clusters=64)
print("fixture", time.time()-t, flush=True)
t=time.time(); ix = f.mirror(); print("mirror", time.time()-t, flush=True)
print(ix.list_sizes()[:8], flush=True)
t=time.time(); cd,k = ix.coarse(f.xq, 16); print("coarse", time.time()-t, cd[0,:4], k[0,:4], flush=True)
cdr,kr = f.ref.coarse(f.xq,16); print(cdr[0,:4], kr[0,:4], flush=True)
t=time.time(); rc,D,I = ix.Search(f.xq, 10, nprobe=8, recall_num=50, metric="L2", has_rank=False); print("search", rc, time.time()-t, D[0,:4], I[0,:4], flush=True)
Dr,Ir = f.ref.search(f.xq,10,json.dumps({"nprobe":8,"recall_num":50,"metric_type":"L2"}),has_rank=False); print(Dr[0,:4], Ir[0,:4], flush=True)
rc,D,I = ix.Search(f.xq, 10, nprobe=8, recall_num=50, metric="L2", has_rank=True); print("search rank", rc, D[0,:4], I[0,:4], flush=True)
Dr,Ir = f.ref.search(f.xq,10,json.dumps({"nprobe":8,"recall_num":50,"metric_type":"L2"}),has_rank=True); print(Dr[0,:4], Ir[0,:4], flush=True)
PY
echo done >> gpurun_out/probe.log
