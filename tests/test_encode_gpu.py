"""GPU encode path: gb200_ivfpq_encode / gb200_ivfpq_add_raw against the reference's own Add
(index/impl/gamma_index_ivfpq.cc:424-512: quantizer->assign + compute_residuals + pq.compute_codes).

Bar: where the coarse assignment agrees, the PQ codes are BIT-IDENTICAL (the kernel reproduces faiss' fvec_L2sqr_ny
arithmetic per slice width); an assignment may differ only when the two centroids are at rounding distance from the
vector (the reference's assign goes through sgemm, whose summation order is the BLAS library's)."""
import json

import numpy as np
import pytest

from conftest import RefFixture, assert_rerank_parity, get_ref_fixture

pytestmark = pytest.mark.gpu


def reference_assignment(f):
    """per vid: (list, code) as the reference's Add produced them"""
    lst = np.full(f.N, -1, np.int32)
    codes = np.zeros((f.N, f.M), np.uint8)
    for l, (ids, cds) in enumerate(f.lists):
        lst[ids] = l
        codes[ids] = cds
    assert (lst >= 0).all()
    return lst, codes


def check_against_reference(f, ln, cd):
    ref_l, ref_c = reference_assignment(f)
    same = ln == ref_l
    # codes: bit-identical wherever the residual is taken against the same centroid
    assert np.array_equal(cd[same], ref_c[same]), "PQ codes differ under the same coarse assignment"
    diff = np.nonzero(~same)[0]
    assert diff.size <= max(2, f.N // 2000), "too many assignment differences: %d" % diff.size
    for v in diff:  # every difference is a rounding-level tie between the two centroids
        x = f.xb[v].astype(np.float64)
        d_ours = ((x - f.centroids[ln[v]].astype(np.float64)) ** 2).sum()
        d_ref = ((x - f.centroids[ref_l[v]].astype(np.float64)) ** 2).sum()
        assert abs(d_ours - d_ref) <= 1e-5 * max(d_ours, d_ref, 1e-12), (v, d_ours, d_ref)
    return float(same.mean())


@pytest.mark.parametrize("d,M", [(128, 32), (128, 64), (128, 16), (96, 8), (96, 16), (64, 64)])
def test_encode_matches_reference_add(d, M):
    """slice widths 4, 2, 8, 12, 6 (generic AVX order) and 1"""
    f = get_ref_fixture("enc_%d_%d" % (d, M), N=12000, d=d, nlist=32, M=M, metric="L2", nq=16, n_clusters=32,
                        seed_shift=40 + M)
    from gamma_b200 import api
    ix = api.B200IVFPQ(0)
    assert ix.Init(f.model_json, f.d) == 0
    ix.set_quantizers(f.centroids, f.pq)
    ln, cd = ix.encode(f.xb)
    agree = check_against_reference(f, ln, cd)
    assert agree > 0.999


def test_add_raw_builds_the_same_index_as_the_reference():
    f = get_ref_fixture("enc_128_32", N=12000, d=128, nlist=32, M=32, metric="L2", nq=16, n_clusters=32, seed_shift=72)
    from gamma_b200 import api
    ix = api.B200IVFPQ(0)
    assert ix.Init(f.model_json, f.d) == 0
    ix.set_quantizers(f.centroids, f.pq)
    # realtime chunks, as AddRTVecsToIndex feeds them (vector/vector_manager.cc:280-382)
    ln_all, cd_all = [], []
    for s in range(0, f.N, 1000):
        ln, cd = ix.add_raw(f.xb[s:s + 1000], first_vid=s)
        ln_all.append(ln)
        cd_all.append(cd)
    ln, cd = np.concatenate(ln_all), np.concatenate(cd_all)
    check_against_reference(f, ln, cd)
    ref_l, _ = reference_assignment(f)
    moved = set(np.nonzero(ln != ref_l)[0].tolist())
    for l in range(f.nlist):  # list order = arrival order, identical to the reference's lists
        ids, codes = ix.get_list(l)
        rids, rcodes = f.lists[l]
        keep = np.array([int(i) not in moved for i in ids], bool)
        rkeep = np.array([int(i) not in moved for i in rids], bool)
        assert np.array_equal(ids[keep], rids[rkeep]) and np.array_equal(codes[keep], rcodes[rkeep])
    rc, D, I = ix.Search(f.xq, 10, nprobe=8, recall_num=60, metric="L2", has_rank=True)
    assert rc == 0
    if not moved:
        assert_rerank_parity(f, ix, f.xq, 10, 8, 60, "L2", D, I)
    else:
        D_ref, I_ref = f.ref.search(f.xq, 10, json.dumps({"nprobe": 8, "recall_num": 60, "metric_type": "L2"}), has_rank=True)
        assert (I == I_ref).mean() > 0.99
