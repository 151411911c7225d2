"""Realtime list maintenance on the device against the reference's own realtime lists:
device compaction (RealTimeMemData::CompactIfNeed / CompactBucket, realtime/realtime_mem_data.cc:354-424, 119-147),
whole-list replacement (what the RetrievalModel plugin uses after Update / compaction on the host), and
searches running concurrently with one another and with a writer (SURVEY §8b threading, tests/test.h:1033-1062)."""
import threading

import numpy as np
import pytest

from conftest import assert_rerank_parity, get_ref_fixture

pytestmark = pytest.mark.gpu

DEL_MASK = np.int64(-2 ** 63)


def own_fixture(seed_shift):
    # a fixture of its own per test (these tests mutate the reference index); kept in the cache like every other one —
    # the reference engine is never torn down inside the test process
    return get_ref_fixture("realtime_%d" % seed_shift, N=20000, d=64, nlist=32, M=16, metric="L2", nq=32, n_clusters=32,
                           seed_shift=seed_shift)


def locate(lists, vid):
    for l, (ids, codes) in enumerate(lists):
        hit = np.nonzero(ids == vid)[0]
        if hit.size:
            return l, int(hit[0]), codes[hit[0]]
    raise AssertionError("vid %d is not alive in any list" % vid)


def test_device_compaction_matches_compact_bucket():
    f = own_fixture(5)
    ix = f.mirror()
    L = int(np.argmax([len(i) for i, _ in f.lists]))
    ids_L = f.lists[L][0]
    kill = ids_L[::2][: int(0.45 * len(ids_L))]  # >= 30 % of the bucket: Compactable (realtime_mem_data.cc:372-376)
    f.delete(kill)
    ix.set_deleted(kill.astype(np.int64), True)  # raises on error
    # an Update anywhere makes the reference run CompactIfNeed (gamma_index_ivfpq.cc:419-421)
    other = (L + 1) % f.nlist
    u = int(f.lists[other][0][0])
    xnew = (f.xb[int(kill[1])] * 1.01).astype(np.float32)
    assert f.ref.update(u, xnew) == 0
    ref_lists = f.ref.lists()
    assert len(ref_lists[L][0]) <= len(ids_L) - len(kill) + 1  # the reference compacted bucket L (u itself may land in it)
    l_new, _, code_new = locate(ref_lists, u)
    assert ix.update(u, l_new, code_new) == 0
    ix.upload_raw(xnew[None, :], first_vid=u)
    dropped = ix.compact(L)
    assert dropped >= len(kill)
    for l in range(f.nlist):
        ids, codes = ix.get_list(l)
        assert np.array_equal(ids, ref_lists[l][0]), "list %d ids differ after compaction" % l
        assert np.array_equal(codes, ref_lists[l][1]), "list %d codes differ after compaction" % l
    f.xb[u] = xnew
    rc, D, I = ix.Search(f.xq, 10, nprobe=8, recall_num=60, metric="L2", has_rank=True)
    assert rc == 0 and not np.isin(I, kill).any()
    assert assert_rerank_parity(f, ix, f.xq, 10, 8, 60, "L2", D, I) > 0.99
    # full compaction: every list loses its dead slots and deleted docs, order kept; results unchanged
    total_before = int(ix.list_sizes().sum())
    dropped_all = ix.compact(-1)
    assert int(ix.list_sizes().sum()) == total_before - dropped_all
    dead = set(int(x) for x in kill)
    for l in range(f.nlist):
        rids, rcodes = ref_lists[l]
        keep = np.array([(i >= 0) and (int(i) not in dead) for i in rids], bool)
        ids, codes = ix.get_list(l)
        assert np.array_equal(ids, rids[keep]) and np.array_equal(codes, rcodes[keep])
    rc, D2, I2 = ix.Search(f.xq, 10, nprobe=8, recall_num=60, metric="L2", has_rank=True)
    assert rc == 0 and np.array_equal(I2, I) and np.array_equal(D2, D)
    # appends keep working after the pools were rebuilt
    v_new = f.N
    assert ix.append(np.array([L], np.int32), np.array([v_new], np.int64), ref_lists[L][1][:1]) == 0
    ids, _ = ix.get_list(L)
    assert ids[-1] == v_new


def test_replace_list_brings_the_device_copy_in_line():
    f = own_fixture(6)
    ix = f.mirror()
    rc, D0, I0 = ix.Search(f.xq, 10, nprobe=8, recall_num=60, metric="L2", has_rank=True)
    assert rc == 0
    L = 7
    ids, codes = f.lists[L]
    assert len(ids) > 40
    # content as the host would hold it after an Update moved one posting away and a compaction dropped a few
    ids2, codes2 = ids.copy(), codes.copy()
    ids2[3] |= DEL_MASK
    keep = np.ones(len(ids2), bool)
    keep[10:20] = False
    ids2, codes2 = ids2[keep], codes2[keep]
    assert ix.replace_list(L, ids2, codes2) == 0
    got_ids, got_codes = ix.get_list(L)
    assert np.array_equal(got_ids, ids2) and np.array_equal(got_codes, codes2)
    gone = set(int(x) for x in ids[10:20]) | {int(ids[3])}
    rc, D1, I1 = ix.Search(f.xq, 10, nprobe=f.nlist, recall_num=60, metric="L2", has_rank=True)
    assert rc == 0 and not (set(I1.ravel().tolist()) & gone)
    # a vid that shows up alive in another list dies where it was (one place per vid)
    mover = int(ids[30])
    other = (L + 5) % f.nlist
    oids, ocodes = f.lists[other]
    assert ix.replace_list(other, np.append(oids, mover), np.vstack([ocodes, codes[30:31]])) == 0
    cur, _ = ix.get_list(L)
    assert (cur[(cur & 0x7fffffff) == mover] < 0).all()
    # putting the original lists back restores the original results bit for bit
    assert ix.replace_list(other, oids, ocodes) == 0
    assert ix.replace_list(L, ids, codes) == 0
    rc, D2, I2 = ix.Search(f.xq, 10, nprobe=8, recall_num=60, metric="L2", has_rank=True)
    assert rc == 0 and np.array_equal(I2, I0) and np.array_equal(D2, D0)
    assert ix.replace_list(L, ids[:0], codes[:0]) == 0  # empty list
    assert ix.get_list(L)[0].size == 0


def test_concurrent_searches_and_a_writer():
    """8 threads search while one appends the second half of the index and another deletes: no call fails, results of
    searches that ran before / after the writer equal the serial ones, and every id returned in between is a posting that
    was visible at some point (never garbage)."""
    f = own_fixture(7)
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
    assert ix.append(list_no[:half], vids[:half], codes[:half]) == 0
    kw = dict(nprobe=8, recall_num=50, metric="L2", has_rank=True)
    rc, D_half, I_half = ix.Search(f.xq, 10, **kw)
    assert rc == 0
    # serial reference of the concurrent part: 8 callers x the same batch
    errors, results = [], [[] for _ in range(8)]
    stop = threading.Event()

    def searcher(t):
        try:
            while not stop.is_set():
                rc_, D_, I_ = ix.Search(f.xq, 10, **kw)
                if rc_ != 0:
                    errors.append(("search rc", rc_, api.lib().gb200_last_error()))
                    return
                results[t].append(I_)
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    def writer():
        try:
            for s in range(half, f.N, 500):
                rc_ = ix.append(list_no[s:s + 500], vids[s:s + 500], codes[s:s + 500])
                if rc_ != 0:
                    errors.append(("append failed", rc_, api.lib().gb200_last_error()))
                    return
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    doomed = np.unique(I_half[:, :3].ravel())[:60].astype(np.int64)

    def deleter():
        try:
            for doc in doomed:
                ix.set_deleted(np.array([doc], np.int64), True)  # raises on error
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=searcher, args=(t,)) for t in range(8)]
    for th in threads:
        th.start()
    w = threading.Thread(target=writer)
    dl = threading.Thread(target=deleter)
    w.start()
    dl.start()
    w.join()
    dl.join()
    stop.set()
    for th in threads:
        th.join()
    assert not errors, errors
    assert sum(len(r) for r in results) >= 8
    for r in results:
        for I_ in r:
            assert I_.min() >= -1 and I_.max() < f.N
    # after the writers: every thread's next search sees the full index without the deleted docs, identical to a serial
    # call and to the reference
    f.delete(doomed)
    rc, D_full, I_full = ix.Search(f.xq, 10, **kw)
    assert rc == 0 and not np.isin(I_full, doomed).any()
    assert_rerank_parity(f, ix, f.xq, 10, 8, 50, "L2", D_full, I_full)
    outs = [None] * 8

    def once(t):
        outs[t] = ix.Search(f.xq, 10, **kw)

    threads = [threading.Thread(target=once, args=(t,)) for t in range(8)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    for rc_, D_, I_ in outs:
        assert rc_ == 0 and np.array_equal(I_, I_full) and np.array_equal(D_, D_full)


def test_merged_batches_equal_serial_calls():
    """Concurrent callers with equal parameters travel in one device batch (capi.cu search_coalesced); callers with
    different parameters do not.  Every caller's rows equal the rows of its own serial call, bit for bit."""
    f = own_fixture(9)
    ix = f.mirror()
    variants = [dict(k=10, nprobe=8, recall_num=50), dict(k=7, nprobe=4, recall_num=30), dict(k=10, nprobe=8, recall_num=50, has_rank=False)]
    plans = []
    for t in range(12):
        v = dict(variants[t % 3])
        lo, hi = (t * 5) % 16, (t * 5) % 16 + 3 + t  # ragged batch sizes 3..14
        plans.append((v, f.xq[lo:hi]))
    serial = []
    for v, xq in plans:
        v = dict(v)
        rc, D, I = ix.Search(xq, v.pop("k"), metric="L2", has_rank=v.pop("has_rank", True), **v)
        assert rc == 0
        serial.append((D, I))
    errors = []

    def caller(t):
        try:
            v, xq = plans[t]
            for _ in range(60):
                vv = dict(v)
                rc, D, I = ix.Search(xq, vv.pop("k"), metric="L2", has_rank=vv.pop("has_rank", True), **vv)
                if rc != 0 or not np.array_equal(I, serial[t][1]) or not np.array_equal(D, serial[t][0]):
                    from gamma_b200 import api
                    errors.append((t, rc, api.lib().gb200_last_error()))
                    return
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=caller, args=(t,)) for t in range(12)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
