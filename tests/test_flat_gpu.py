"""GPU parity tests of the FLAT scan (GammaFLATIndex::Search) against the compiled reference."""
import json

import numpy as np
import pytest

from conftest import assert_topk_parity

pytestmark = pytest.mark.gpu
FLT_MAX = float(np.finfo(np.float32).max)


def build(N, d, metric, nq, normalize=False):
    from gamma_b200 import api, synth
    from oracle import ref
    xb = synth.mixture(N, d, synth.SEED_BASE + 11, n_clusters=64, normalize=normalize)
    xq = synth.mixture(nq, d, synth.SEED_QUERY + 11, n_clusters=64, normalize=normalize)
    r = ref.RefIndex(d, "FLAT", json.dumps({"metric_type": metric}), indexing_size=N, bitmap_bits=max(2 * N, 1024))
    r.add_raw(xb)
    ix = api.B200FLAT(0)
    assert ix.Init(json.dumps({"metric_type": metric}), d) == 0
    assert ix.Add(xb)
    return xb, xq, r, ix


def test_flat_l2_config1_smoke_shape():
    """BASELINE config[0]: FLAT L2 d=64, 100k vectors, batch=1, k=10."""
    xb, xq, r, ix = build(100000, 64, "L2", 4)
    for i in range(4):
        D_ref, I_ref = r.search(xq[i:i + 1], 10, json.dumps({"metric_type": "L2"}))
        rc, D, I = ix.Search(xq[i:i + 1], 10, metric="L2")
        assert rc == 0
        assert np.array_equal(I_ref, I) and np.array_equal(D_ref, D)  # bit-exact ids AND distances


@pytest.mark.parametrize("metric,d", [("L2", 48), ("InnerProduct", 100), ("L2", 7)])
def test_flat_batch_filters_window(metric, d):
    from gamma_b200 import synth
    N = 20000
    xb, xq, r, ix = build(N, d, metric, 16, normalize=(metric != "L2"))
    field = synth.filter_field(N)
    flags = (field < 30).astype(np.uint8)
    dele = synth.deleted_docs(N, 0.01)
    for doc in dele:
        r.delete(int(doc))
    ix.set_deleted(dele)
    filt = [(0, N - 1, False, flags)]
    pj = json.dumps({"metric_type": metric, "parallel_on_queries": 0})
    D_ref, I_ref = r.search(xq, 10, pj, filters=filt)
    rc, D, I = ix.Search(xq, 10, metric=metric, filters=filt)
    assert rc == 0
    assert np.all(flags[I[I >= 0]] == 1) and not np.isin(I[I >= 0], dele).any()
    assert_topk_parity(D_ref, I_ref, D, I, rtol=1e-6, atol=0.0)
    assert (D_ref == D).mean() > 0.99  # same summation order as the AVX kernels
    # score window (applied before the heap on the flat path)
    lo, hi = float(np.percentile(D_ref[:, 2], 50)), float(np.percentile(D_ref[:, 8], 50))
    lo, hi = min(lo, hi), max(lo, hi)
    D_ref, I_ref = r.search(xq, 10, pj, filters=filt, min_score=lo, max_score=hi)
    rc, D, I = ix.Search(xq, 10, metric=metric, filters=filt, min_score=lo, max_score=hi)
    assert rc == 0
    assert_topk_parity(D_ref, I_ref, D, I, rtol=1e-6, atol=0.0)
    neutral = FLT_MAX if metric == "L2" else -FLT_MAX
    assert np.all(D[I < 0] == neutral)


def test_flat_empty_store_and_k_gt_n():
    from gamma_b200 import api
    ix = api.B200FLAT(0)
    assert ix.Init('{"metric_type": "L2"}', 16) == 0
    xq = np.zeros((2, 16), np.float32)
    rc, D, I = ix.Search(xq, 5, metric="L2")
    assert rc == 0 and np.all(I == -1) and np.all(D == FLT_MAX)
    ix.Add(np.eye(3, 16, dtype=np.float32))
    rc, D, I = ix.Search(xq, 5, metric="L2")
    assert rc == 0
    assert np.array_equal(I[0], [0, 1, 2, -1, -1]) and np.all(D[0, :3] == 1.0)


@pytest.mark.parametrize("metric,d,N,nq", [("InnerProduct", 64, 150000, 64), ("L2", 96, 70000, 40),
                                           ("InnerProduct", 768, 30000, 48)])  # d = 768: BASELINE config 4 (24 k-blocks)
def test_flat_batch_tensor_core_path_is_exact(metric, d, N, nq, monkeypatch):
    """Batches (n >= 16) take the tcgen05 GEMM -> candidate select -> exact re-score route; the result
    must still be bit-identical to the CPU engine (ids and distances), including across database chunks."""
    xb, xq, r, ix = build(N, d, metric, nq, normalize=(metric != "L2"))
    pj = json.dumps({"metric_type": metric, "parallel_on_queries": 0})
    D_ref, I_ref = r.search(xq, 10, pj)
    rc, D, I = ix.Search(xq, 10, metric=metric)
    assert rc == 0
    assert np.array_equal(I_ref, I) and np.array_equal(D_ref, D)
    # many database chunks: the running candidate state is carried (and its threshold applied) across chunk selects;
    # a filter + a score window ride along; the older select kernel must agree
    flags = (np.arange(N) % 3 != 0).astype(np.uint8)
    filt = [(0, N - 1, False, flags)]
    D_reff, I_reff = r.search(xq, 10, pj, filters=filt)
    for env in ({"GB200_FLAT_CHUNK_ROWS": "4096"}, {"GB200_FLAT_CHUNK_ROWS": "8192", "GB200_FLAT_SELECT_V1": "1"}):
        for kk, vv in env.items():
            monkeypatch.setenv(kk, vv)
        ix.reload_tuning()
        rc, D3, I3 = ix.Search(xq, 10, metric=metric)
        assert rc == 0 and np.array_equal(I_ref, I3) and np.array_equal(D_ref, D3), env
        rc, D4, I4 = ix.Search(xq, 10, metric=metric, filters=filt)
        assert rc == 0 and np.array_equal(I_reff, I4) and np.array_equal(D_reff, D4), env
        for kk in env:
            monkeypatch.delenv(kk)
    # the per-query exact scan (GB200_FLAT=exact) gives the same answer
    monkeypatch.setenv("GB200_FLAT", "exact")
    ix.reload_tuning()
    rc, D2, I2 = ix.Search(xq, 10, metric=metric)
    assert np.array_equal(I, I2) and np.array_equal(D, D2)
