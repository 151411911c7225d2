"""GPU parity tests of the IVFPQ hot path against the compiled reference engine (oracle/_ref):
same trained quantizers, same postings in the same list order, same queries, same filters.

Parity contract: ids identical except where the reference ordering is a
floating-point near-tie; ADC distances within 1e-4 relative; re-ranked (exact) distances
bit-identical because exact_distance reproduces faiss' AVX summation order.
"""
import json
import os

import numpy as np
import pytest

from conftest import assert_rerank_parity, assert_topk_parity, compare_topk, get_ref_fixture

pytestmark = pytest.mark.gpu

FLT_MAX = float(np.finfo(np.float32).max)


def fx_l2_m32():
    return get_ref_fixture("l2_m32", N=40000, d=128, nlist=128, M=32, metric="L2", nq=96, n_clusters=128)


def fx_l2_m16():
    return get_ref_fixture("l2_m16", N=20000, d=64, nlist=64, M=16, metric="L2", nq=48, n_clusters=64)


def fx_ip_m32():
    return get_ref_fixture("ip_m32", N=30000, d=128, nlist=96, M=32, metric="InnerProduct", nq=64, n_clusters=96)


def rj(nprobe, recall_num, metric):
    return json.dumps({"nprobe": nprobe, "recall_num": recall_num, "metric_type": metric})


@pytest.mark.parametrize("fx", [fx_l2_m32, fx_l2_m16, fx_ip_m32])
def test_posting_mirror_roundtrip(fx):
    f = fx()
    ix = f.mirror(raw=False)
    sizes = ix.list_sizes()
    assert int(sizes.sum()) == f.N
    for l in [0, 1, f.nlist // 2, f.nlist - 1]:
        ids, codes = ix.get_list(l)
        rids, rcodes = f.lists[l]
        assert np.array_equal(ids, rids)  # bit-exact list order
        assert np.array_equal(codes, rcodes)  # bit-exact codes through the blocked/rotated layout


@pytest.mark.parametrize("fx", [fx_l2_m32, fx_l2_m16])
def test_coarse_quantizer_matches_reference(fx):
    f = fx()
    ix = f.mirror(raw=False)
    nprobe = 16
    cd_ref, k_ref = f.ref.coarse(f.xq, nprobe)
    cd, k = ix.coarse(f.xq, nprobe)
    r = compare_topk(cd_ref, k_ref, cd, k, rtol=1e-4, atol=1e-4)
    assert r["n_id_mismatch_unexplained"] == 0, r
    assert r["max_rel_err"] <= 1e-4, r
    assert np.all(np.diff(cd, axis=1) >= 0)  # ascending coarse distance


@pytest.mark.parametrize("fx,metric", [(fx_l2_m32, "L2"), (fx_l2_m16, "L2"), (fx_ip_m32, "InnerProduct")])
def test_adc_scan_preassigned_no_rank(fx, metric):
    """ADC scan + recall selection in isolation: probes taken from the reference's own coarse stage
    (search_preassigned), has_rank off so returned distances are the ADC sums."""
    f = fx()
    ix = f.mirror()
    nprobe, R = 12, 50
    cd_ref, k_ref = f.ref.coarse(f.xq, nprobe)
    D_ref, I_ref = f.ref.search(f.xq, R, rj(nprobe, R, metric), has_rank=False, keys=k_ref, coarse_dis=cd_ref)
    rc, D, I = ix.Search(f.xq, R, nprobe=nprobe, recall_num=R, metric=metric, has_rank=False, keys=k_ref,
                         coarse_dis=cd_ref)
    assert rc == 0
    r = assert_topk_parity(D_ref, I_ref, D, I, rtol=1e-4, atol=1e-5)
    assert ix.last_scanned_postings() > 0


@pytest.mark.parametrize("fx,metric", [(fx_l2_m32, "L2"), (fx_l2_m16, "L2"), (fx_ip_m32, "InnerProduct")])
def test_full_search_with_rerank(fx, metric):
    f = fx()
    ix = f.mirror()
    nprobe, R, k = 16, 100, 10
    rc, D, I = ix.Search(f.xq, k, nprobe=nprobe, recall_num=R, metric=metric, has_rank=True)
    assert rc == 0
    # exact re-rank reproduces the AVX summation order: bit-identical distances where ids agree; ids differ only at
    # the boundary of the recall set (see assert_rerank_parity)
    assert assert_rerank_parity(f, ix, f.xq, k, nprobe, R, metric, D, I) > 0.995


def test_large_batch_work_plan_parity(monkeypatch):
    """More queries than resident CTA slots (SMs x 3): the persistent scan hands queries to CTAs from a queue and lets
    idle CTAs join running queries (several candidate rows per query, merged by the re-rank) — it must still
    reproduce the CPU engine, and every other way of cutting the batch (one row per query, tiny items, other CTA shapes,
    the bulk-copy fed posting ring) must return bit-identical results."""
    from gamma_b200 import synth
    f = fx_l2_m32()
    ix = f.mirror()
    nprobe, R, k = 16, 100, 10
    xq = synth.mixture(1100, f.d, synth.SEED_QUERY + 5, n_clusters=128)
    rc, D, I = ix.Search(xq, k, nprobe=nprobe, recall_num=R, metric="L2", has_rank=True)
    assert rc == 0
    assert assert_rerank_parity(f, ix, xq, k, nprobe, R, "L2", D, I) > 0.995
    for env in ({"GB200_SCAN_ROWS": "1"}, {"GB200_SCAN_CH": "1", "GB200_SCAN_HELP_MIN": "1", "GB200_SCAN_ROWS": "8"},
                {"GB200_SCAN_THREADS": "512"}, {"GB200_SCAN_THREADS": "320"}, {"GB200_SCAN_THREADS": "256", "GB200_SCAN_CH": "3"},
                {"GB200_SCAN_TMA": "1"}, {"GB200_SCAN_TMA": "1", "GB200_SCAN_THREADS": "256", "GB200_SCAN_CH": "2"}):
        for kk, vv in env.items():
            monkeypatch.setenv(kk, vv)
        ix.reload_tuning()
        for _ in range(3):  # which CTA scans what is timing dependent; the result must not be
            rc, D2, I2 = ix.Search(xq, k, nprobe=nprobe, recall_num=R, metric="L2", has_rank=True)
            assert rc == 0 and np.array_equal(I2, I) and np.array_equal(D2, D), env
        for kk in env:
            monkeypatch.delenv(kk)
    ix.reload_tuning()
    # has_rank = False: the ADC distances themselves, merged across the rows of a query
    D_ref, I_ref = f.ref.search(xq, k, rj(nprobe, R, "L2"), has_rank=False)
    rc, D, I = ix.Search(xq, k, nprobe=nprobe, recall_num=R, metric="L2", has_rank=False)
    assert rc == 0
    assert_topk_parity(D_ref, I_ref, D, I, rtol=1e-4, atol=1e-5)


def fx_l2_nlist4096():
    return get_ref_fixture("l2_nlist4096", N=200000, d=64, nlist=4096, M=32, metric="L2", nq=8, n_clusters=512)


@pytest.mark.parametrize("n", [4096, 4097])
def test_many_lists_nprobe64_batch_4096(n):
    """The shape class of BASELINE config 5 at test size: nlist = 4096, nprobe = 64 (two 32-wide groups in every probe
    table walk), batch 4096 / 4097 (beyond the old work-plan limit), short lists (49 postings on average)."""
    from gamma_b200 import synth
    f = fx_l2_nlist4096()
    ix = f.mirror()
    nprobe, R, k = 64, 100, 10
    xq = synth.mixture(n, f.d, synth.SEED_QUERY + 21, n_clusters=512)
    cd_ref, k_ref = f.ref.coarse(xq, nprobe)
    cd, kk = ix.coarse(xq, nprobe)
    r = compare_topk(cd_ref, k_ref, cd, kk, rtol=1e-4, atol=1e-4)
    assert r["n_id_mismatch_unexplained"] == 0 and r["max_rel_err"] <= 1e-4, r
    rc, D, I = ix.Search(xq, k, nprobe=nprobe, recall_num=R, metric="L2", has_rank=True)
    assert rc == 0
    assert assert_rerank_parity(f, ix, xq, k, nprobe, R, "L2", D, I) > 0.995
    D_ref, I_ref = f.ref.search(xq, k, rj(nprobe, R, "L2"), has_rank=False, keys=k_ref, coarse_dis=cd_ref)
    rc, D, I = ix.Search(xq, k, nprobe=nprobe, recall_num=R, metric="L2", has_rank=False, keys=k_ref, coarse_dis=cd_ref)
    assert rc == 0
    assert_topk_parity(D_ref, I_ref, D, I, rtol=1e-4, atol=1e-5)


def test_filters_and_deletions_inside_the_scan():
    from gamma_b200 import synth
    f = fx_l2_m32()
    ix = f.mirror()
    N = f.N
    field = synth.filter_field(N)
    pass_flags = (field < 30).astype(np.uint8)
    dele = synth.deleted_docs(N, 0.01)
    f.delete(dele)  # recorded on the shared fixture: every later mirror() replays the same deletions
    try:
        ix.set_deleted(dele, True)
        filt = [(0, N - 1, False, pass_flags)]
        nprobe, R, k = 16, 100, 10
        rc, D, I = ix.Search(f.xq, k, nprobe=nprobe, recall_num=R, metric="L2", has_rank=True, filters=filt)
        assert rc == 0
        got = I[I >= 0]
        assert np.all(pass_flags[got] == 1) and not np.isin(got, dele).any()
        assert_rerank_parity(f, ix, f.xq, k, nprobe, R, "L2", D, I, filters=filt)
        # a second, partial-range NOT-IN filter on top (b_not_in_ and min/max clipping semantics)
        lo, hi = 1003, N // 2 + 5
        flags2 = (np.arange(lo, hi + 1) % 3 == 0).astype(np.uint8)
        filt2 = filt + [(lo, hi, True, flags2)]
        D_ref, I_ref = f.ref.search(f.xq, k, rj(nprobe, R, "L2"), has_rank=False, filters=filt2)
        rc, D, I = ix.Search(f.xq, k, nprobe=nprobe, recall_num=R, metric="L2", has_rank=False, filters=filt2)
        assert rc == 0
        assert_topk_parity(D_ref, I_ref, D, I, rtol=1e-4, atol=1e-5)
        # deletions only (no range filter)
        rc, D, I = ix.Search(f.xq, k, nprobe=nprobe, recall_num=R, metric="L2", has_rank=True)
        assert not np.isin(I[I >= 0], dele).any()
        assert_rerank_parity(f, ix, f.xq, k, nprobe, R, "L2", D, I)
    finally:
        pass  # fixture is cached with the deletions applied; later tests re-apply the same set


def test_score_window_and_unfilled_slots():
    f = fx_l2_m16()
    ix = f.mirror()
    nprobe, R, k = 8, 40, 20
    D0, _ = f.ref.search(f.xq, k, rj(nprobe, R, "L2"), has_rank=True)
    lo, hi = float(np.percentile(D0[:, 0], 50)), float(np.percentile(D0[:, 5], 50))
    for has_rank in (True, False):
        D_ref, I_ref = f.ref.search(f.xq, k, rj(nprobe, R, "L2"), has_rank=has_rank, min_score=lo, max_score=hi)
        rc, D, I = ix.Search(f.xq, k, nprobe=nprobe, recall_num=R, metric="L2", has_rank=has_rank, min_score=lo,
                             max_score=hi)
        assert rc == 0
        assert (I == -1).any()  # the window leaves unfilled slots
        filled = I >= 0
        assert np.array_equal(filled, I_ref >= 0) or compare_topk(D_ref, I_ref, D, I)["n_id_mismatch_unexplained"] == 0
        assert np.all(D[~filled] == FLT_MAX)  # heap neutral value for L2
        assert_topk_parity(D_ref, I_ref, D, I, rtol=1e-4 if not has_rank else 1e-6, atol=1e-5 if not has_rank else 0.0)


def test_k_larger_than_recall_and_few_candidates():
    f = fx_l2_m16()
    ix = f.mirror()
    # recall_num < k  => recall_num = k (gamma_index_ivfpq.cc:762-765); nprobe=1 => fewer than k candidates for some
    D_ref, I_ref = f.ref.search(f.xq, 64, rj(1, 10, "L2"), has_rank=True)
    rc, D, I = ix.Search(f.xq, 64, nprobe=1, recall_num=10, metric="L2", has_rank=True)
    assert rc == 0
    assert_topk_parity(D_ref, I_ref, D, I, rtol=1e-6, atol=0.0)


def test_error_codes_match_reference_convention():
    f = fx_l2_m16()
    ix = f.mirror(raw=False)
    rc, _, _ = ix.Search(f.xq, 10, nprobe=4, recall_num=20, has_rank=True)  # has_rank without raw vectors
    assert rc < 0
    from gamma_b200 import api
    fresh = api.B200IVFPQ(0)
    assert fresh.Init(f.model_json, f.d) == 0
    rc, _, _ = fresh.Search(f.xq, 10)
    assert rc == -2  # untrained: reference GammaEngine::Search returns -2 when not indexed


def test_realtime_append_and_update_visible_to_next_search():
    """Postings appended after a search are seen by the next one (RTInvertIndex::AddKeys semantics);
    Update to another list kills the old posting (kDelIdxMask) and appends the new one."""
    f = fx_l2_m16()
    from gamma_b200 import api
    ix = api.B200IVFPQ(0)
    assert ix.Init(f.model_json, f.d) == 0
    ix.set_quantizers(f.centroids, f.pq)
    list_no = np.concatenate([np.full(len(ids), l, np.int32) for l, (ids, _) in enumerate(f.lists)])
    vids = np.concatenate([ids for ids, _ in f.lists])
    codes = np.concatenate([c for _, c in f.lists])
    order = np.argsort(vids, kind="stable")
    list_no, vids, codes = list_no[order], vids[order], codes[order]
    half = f.N // 2
    ix.upload_raw(f.xb)
    # chunks of 1000 like AddRTVecsToIndex (vector_manager.cc:280-382)
    for s in range(0, half, 1000):
        assert ix.append(list_no[s:s + 1000], vids[s:s + 1000], codes[s:s + 1000]) == 0
    rc, D1, I1 = ix.Search(f.xq, 10, nprobe=8, recall_num=50, metric="L2", has_rank=True)
    assert rc == 0 and I1.max() < half
    for s in range(half, f.N, 1000):
        assert ix.append(list_no[s:s + 1000], vids[s:s + 1000], codes[s:s + 1000]) == 0
    rc, D2, I2 = ix.Search(f.xq, 10, nprobe=8, recall_num=50, metric="L2", has_rank=True)
    assert rc == 0
    assert_rerank_parity(f, ix, f.xq, 10, 8, 50, "L2", D2, I2)
    for l in [0, f.nlist - 1]:
        ids, cds = ix.get_list(l)
        assert np.array_equal(ids, f.lists[l][0]) and np.array_equal(cds, f.lists[l][1])
    # Update: move vid v to another list with a new code
    v = int(I2[0, 0])
    old_list = int(list_no[np.nonzero(vids == v)[0][0]])
    new_list = (old_list + 1) % f.nlist
    new_code = codes[np.nonzero(vids == v)[0][0]].copy()
    assert ix.update(v, new_list, new_code) == 0
    ids_old, _ = ix.get_list(old_list)
    dead = ids_old[(ids_old & 0x7fffffff) == v]
    assert dead.size == 1 and dead[0] < 0  # kDelIdxMask bit set
    ids_new, codes_new = ix.get_list(new_list)
    assert ids_new[-1] == v and np.array_equal(codes_new[-1], new_code)


@pytest.mark.parametrize("fx,metric", [(fx_l2_m32, "L2"), (fx_l2_m16, "L2")])
def test_large_recall_num_uses_the_wide_select(fx, metric):
    """recall_num = 600 and 1700 (> 512): candidate buffers of 2048 / 4096 keys in the 512-thread shape (4 / 8 keys per
    thread in the radix select)."""
    f = fx()
    ix = f.mirror()
    for nprobe, R, k in ((24, 600, 50), (32, 1700, 20)):
        cd_ref, k_ref = f.ref.coarse(f.xq, nprobe)
        rc, D, I = ix.Search(f.xq, k, nprobe=nprobe, recall_num=R, metric=metric, has_rank=True, keys=k_ref, coarse_dis=cd_ref)
        assert rc == 0
        assert_rerank_parity(f, ix, f.xq, k, nprobe, R, metric, D, I, keys=k_ref, coarse_dis=cd_ref)
        D_ref, I_ref = f.ref.search(f.xq, R, rj(nprobe, R, metric), has_rank=False, keys=k_ref, coarse_dis=cd_ref)
        rc, D, I = ix.Search(f.xq, R, nprobe=nprobe, recall_num=R, metric=metric, has_rank=False, keys=k_ref, coarse_dis=cd_ref)
        assert rc == 0
        assert_topk_parity(D_ref, I_ref, D, I, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("nprobe", [1, 31, 32, 33, 64, 100, 128])
def test_coarse_select_nprobe_sweep(nprobe):
    """warp-per-row select (nprobe <= 128): all KPL variants, ascending order, ties by centroid id."""
    f = fx_l2_m32()
    ix = f.mirror(raw=False)
    cd_ref, k_ref = f.ref.coarse(f.xq, nprobe)
    cd, k = ix.coarse(f.xq, nprobe)
    r = compare_topk(cd_ref, k_ref, cd, k, rtol=1e-4, atol=1e-4)
    assert r["n_id_mismatch_unexplained"] == 0 and r["max_rel_err"] <= 1e-4, r
    assert np.all(np.diff(cd, axis=1) >= 0)


@pytest.mark.parametrize("nprobe", [1, 16, 33, 64, 128])
def test_coarse_select_from_chunk_minima(nprobe, monkeypatch):
    """nlist >= 4096: the select starts from the chunk minima the tensor-core GEMM emits (coarse_select_cmin_kernel).
    Parity with the reference's quantizer->search, and bit-identical to the full-row select of the same distances."""
    from gamma_b200 import synth
    f = fx_l2_nlist4096()
    ix = f.mirror(raw=False)
    xq = synth.mixture(300, f.d, synth.SEED_QUERY + 33, n_clusters=512)
    xq[7] = 0.0  # all-zero query
    cd_ref, k_ref = f.ref.coarse(xq, nprobe)
    cd, k = ix.coarse(xq, nprobe)
    r = compare_topk(cd_ref, k_ref, cd, k, rtol=1e-4, atol=1e-4)
    assert r["n_id_mismatch_unexplained"] == 0 and r["max_rel_err"] <= 1e-4, r
    assert np.all(np.diff(cd, axis=1) >= 0)
    xq[8] = f.centroids[5]  # a query that IS a centroid: |q|^2 + |c|^2 - 2 q.c cancels to ~0 (clamped), both paths agree
    cd, k = ix.coarse(xq, nprobe)
    monkeypatch.setenv("GB200_COARSE_FULL_SELECT", "1")
    ix.reload_tuning()
    cd2, k2 = ix.coarse(xq, nprobe)
    assert np.array_equal(k, k2) and np.array_equal(cd, cd2)


@pytest.mark.parametrize("metric", ["L2", "InnerProduct"])
def test_m64_default_kernel_parity(metric):
    """M = 64 is the reference's default nsubvector (gamma_index_ivfpq.h:693) and BASELINE config 3: the conflict-free
    kernel over the pre-rotated layout is the default path."""
    from gamma_b200 import synth
    f = get_ref_fixture("m64_" + metric, N=40000, d=128, nlist=128, M=64, metric=metric, nq=96, n_clusters=128)
    ix = f.mirror()
    for l in [0, 1, f.nlist // 2, f.nlist - 1]:  # the rotated layout reads back as the reference's AoS lists
        ids, codes = ix.get_list(l)
        rids, rcodes = f.lists[l]
        assert np.array_equal(ids, rids) and np.array_equal(codes, rcodes)
    nprobe, R = 12, 50
    cd_ref, k_ref = f.ref.coarse(f.xq, nprobe)
    D_ref, I_ref = f.ref.search(f.xq, R, rj(nprobe, R, metric), has_rank=False, keys=k_ref, coarse_dis=cd_ref)
    rc, D, I = ix.Search(f.xq, R, nprobe=nprobe, recall_num=R, metric=metric, has_rank=False, keys=k_ref, coarse_dis=cd_ref)
    assert rc == 0
    assert_topk_parity(D_ref, I_ref, D, I, rtol=1e-4, atol=1e-5)
    # full search with re-rank, a batch larger than the resident CTA slots (work plan with a split tail), a filter
    normalize = metric != "L2"
    xq = synth.mixture(700, f.d, synth.SEED_QUERY + 9, n_clusters=128, normalize=normalize)
    flags = (synth.filter_field(f.N) < 50).astype(np.uint8)
    for filt in ([], [(0, f.N - 1, False, flags)]):
        rc, D, I = ix.Search(xq, 10, nprobe=16, recall_num=100, metric=metric, has_rank=True, filters=filt)
        assert rc == 0
        assert assert_rerank_parity(f, ix, xq, 10, 16, 100, metric, D, I, filters=filt) > 0.995
    # large recall_num: the 16-keys-per-thread select variant
    rc, D, I = ix.Search(f.xq, 10, nprobe=16, recall_num=700, metric=metric, has_rank=True)
    assert rc == 0
    assert_rerank_parity(f, ix, f.xq, 10, 16, 700, metric, D, I)


def test_opq_model_parity():
    """Model parameter "opq": queries and added vectors go through the OPQ matrix before the quantizers, the re-rank keeps
    the raw query (gamma_index_ivfpq.cc:158-165, 448-450, 547-555, 706).  The reference applies the matrix through sgemm,
    the device with an fp32 FMA chain: ADC distances agree to 1e-4, re-ranked distances are bit-identical, encoded codes
    agree except where the two roundings fall on different sides of a PQ cell boundary."""
    f = get_ref_fixture("l2_m16_opq", N=20000, d=64, nlist=32, M=16, metric="L2", nq=48, n_clusters=32, seed_shift=9,
                        extra_params={"opq": {"nsubvector": 16}})
    assert f.opq is not None
    ix = f.mirror()
    nprobe, R, k = 8, 60, 10
    cd_ref, k_ref = f.ref.coarse(f.xq @ f.opq[0].T + (0 if f.opq[1] is None else f.opq[1]), nprobe)
    cd, kk = ix.coarse(f.xq, nprobe)
    r = compare_topk(cd_ref, k_ref, cd, kk, rtol=1e-4, atol=1e-4)
    assert r["n_id_mismatch_unexplained"] == 0 and r["max_rel_err"] <= 1e-4, r
    D_ref, I_ref = f.ref.search(f.xq, R, rj(nprobe, R, "L2"), has_rank=False)
    rc, D, I = ix.Search(f.xq, R, nprobe=nprobe, recall_num=R, metric="L2", has_rank=False)
    assert rc == 0
    assert_topk_parity(D_ref, I_ref, D, I, rtol=1e-4, atol=1e-5)
    rc, D, I = ix.Search(f.xq, k, nprobe=nprobe, recall_num=R, metric="L2", has_rank=True)
    assert rc == 0
    assert assert_rerank_parity(f, ix, f.xq, k, nprobe, R, "L2", D, I) > 0.99
    # encode through the same transform
    ln, cd_ = ix.encode(f.xb[:4000])
    ref_l = np.full(f.N, -1, np.int32)
    ref_c = np.zeros((f.N, f.M), np.uint8)
    for l, (ids, cds) in enumerate(f.lists):
        ref_l[ids] = l
        ref_c[ids] = cds
    same = ln == ref_l[:4000]
    assert same.mean() > 0.998
    assert (cd_[same] == ref_c[:4000][same]).mean() > 0.999
