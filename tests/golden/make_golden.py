"""Generates tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref = unmodified
GammaIVFPQIndex / GammaFLATIndex over faiss 1.7.1, built by oracle/Makefile from /root/reference).

    python tests/golden/make_golden.py

The fixtures hold inputs (trained quantizers, realtime lists in the reference layout, raw vectors,
queries, filter) and the reference's outputs, so the restatement (CPU tests) and the CUDA path (GPU
tests) can be checked where /root/reference does not exist.  Seeds are fixed; faiss k-means is seeded
(ClusteringParameters.seed = 1234) and training here runs with a fixed thread count."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gamma_b200 import synth  # noqa: E402
from oracle import ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make(name, N, d, nlist, M, metric, nq, nprobe, R):
    normalize = metric != "L2"
    xb = synth.mixture(N, d, 101, n_clusters=nlist, normalize=normalize)
    xq = synth.mixture(nq, d, 102, n_clusters=nlist, normalize=normalize)
    mj = json.dumps({"ncentroids": nlist, "nsubvector": M, "metric_type": metric, "nprobe": nprobe})
    r = ref.RefIndex(d, "IVFPQ", mj, indexing_size=N, bitmap_bits=4 * N)
    r.add_raw(xb)
    r.indexing(threads=4)
    r.add_to_index()
    # one Update that moves a posting to another list (kDelIdxMask on the old one)
    moved_vid = 17
    r.update(moved_vid, xb[N - 1])
    xb[moved_vid] = xb[N - 1]
    lists = r.lists()
    rj = json.dumps({"nprobe": nprobe, "recall_num": R, "metric_type": metric})
    cd, keys = r.coarse(xq, nprobe)
    out = dict(xb=xb, xq=xq, centroids=r.centroids(), pq=r.pq_centroids(), coarse_dis=cd, coarse_keys=keys,
               list_lens=np.array([len(i) for i, _ in lists], np.int64),
               list_ids=np.concatenate([i for i, _ in lists]), list_codes=np.concatenate([c for _, c in lists]),
               meta=np.array(json.dumps(dict(N=N, d=d, nlist=nlist, M=M, metric=metric, nprobe=nprobe, R=R))))
    D, I = r.search(xq, R, rj, has_rank=False)
    out["norank_D"], out["norank_I"] = D, I
    D, I = r.search(xq, 10, rj, has_rank=True)
    out["rank_D"], out["rank_I"] = D, I
    flags = (synth.filter_field(N) < 30).astype(np.uint8)
    dele = synth.deleted_docs(N, 0.02)
    lo, hi = N // 10 + 3, N - N // 7
    flags2 = (np.arange(lo, hi + 1) % 5 != 0).astype(np.uint8)
    for doc in dele:
        r.delete(int(doc))
    filt = [(0, N - 1, False, flags), (lo, hi, True, flags2)]
    D, I = r.search(xq, 10, rj, has_rank=True, filters=filt)
    out.update(filt_flags=flags, filt2_lo=lo, filt2_hi=hi, filt2_flags=flags2, deleted=dele, filt_rank_D=D, filt_rank_I=I)
    D, I = r.search(xq, R, rj, has_rank=False, filters=filt[:1])
    out["filt_norank_D"], out["filt_norank_I"] = D, I
    # score window taken from the un-windowed result
    w_lo, w_hi = float(np.sort(out["rank_D"][:, 1])[nq // 2]), float(np.sort(out["rank_D"][:, 6])[nq // 2])
    w_lo, w_hi = min(w_lo, w_hi), max(w_lo, w_hi)
    D, I = r.search(xq, 10, rj, has_rank=True, filters=filt, min_score=w_lo, max_score=w_hi)
    out.update(window=np.array([w_lo, w_hi], np.float32), win_rank_D=D, win_rank_I=I)
    # FLAT model over the same raw vectors, deletions and filters.  (NOT the IVFPQ model's
    # brute_force_search fallback: gamma_index_ivfpq.cc:531-533 deletes the retrieval params and then
    # reads metric/parallel flags from the freed object, so that path's metric is undefined behaviour.)
    fl = ref.RefIndex(d, "FLAT", json.dumps({"metric_type": metric}), indexing_size=N, bitmap_bits=4 * N)
    fl.add_raw(xb)
    for doc in dele:
        fl.delete(int(doc))
    D, I = fl.search(xq, 10, json.dumps({"metric_type": metric}), filters=filt)
    out["flat_D"], out["flat_I"] = D, I
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: getattr(v, "shape", None) for k, v in out.items() if k.endswith("_I")})


if __name__ == "__main__":
    make("ivfpq_l2_d32_m8", N=3000, d=32, nlist=16, M=8, metric="L2", nq=12, nprobe=4, R=20)
    make("ivfpq_ip_d64_m32", N=4000, d=64, nlist=16, M=32, metric="InnerProduct", nq=12, nprobe=5, R=24)
