import json, os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from conftest import get_ref_fixture, compare_topk
f = get_ref_fixture("l2_m32", N=40000, d=128, nlist=128, M=32, metric="L2", nq=96, n_clusters=128)
nprobe = 24
cd_ref, k_ref = f.ref.coarse(f.xq, nprobe)
for R in (100, 400, 512, 513, 600, 1000):
    rj = json.dumps({"nprobe": nprobe, "recall_num": R, "metric_type": "L2"})
    D_ref, I_ref = f.ref.search(f.xq, R, rj, has_rank=False, keys=k_ref, coarse_dis=cd_ref)
    for env in ({}, {"GB200_SCAN_SPLITS": "1"}, {"GB200_SCAN_FORCE_SYM": "1"}, {"GB200_FORCE_GENERIC": "1"}):
        for k_, v_ in env.items(): os.environ[k_] = v_
        ix = f.mirror()
        rc, D, I = ix.Search(f.xq, R, nprobe=nprobe, recall_num=R, metric="L2", has_rank=False, keys=k_ref, coarse_dis=cd_ref)
        r = compare_topk(D_ref, I_ref, D, I, rtol=1e-4, atol=1e-5)
        badq = int((np.sort(I, 1) != np.sort(I_ref, 1)).any(1).sum())
        nfill = int((I >= 0).sum()), int((I_ref >= 0).sum())
        print("R", R, env, "rc", rc, r, "queries with different id sets", badq, "filled", nfill, flush=True)
        for k_ in env: os.environ.pop(k_)
        ix.close()
