import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def has_gpu():
    try:
        from gamma_b200 import api
        return api.lib().gb200_device_count() > 0
    except Exception:
        return False


# ----------------------------------------------------------------------------------------------
# result comparison: ids must agree except where the reference's own ordering is a floating-point
# (near-)tie; distances must agree to `rtol` relative.
# ----------------------------------------------------------------------------------------------
def compare_topk(D_ref, I_ref, D, I, rtol=1e-4, atol=1e-6):
    """Returns dict(n_id_mismatch_unexplained, max_rel_err, n_slots).  A differing id at slot j is
    'explained' when it is a near-tie swap: the id we returned sits in the reference list at a
    distance within tolerance of slot j's, or (boundary tie) it is absent from the reference list but
    its distance is within tolerance of the reference's last filled slot."""
    D_ref, I_ref, D, I = map(np.asarray, (D_ref, I_ref, D, I))
    n, k = I_ref.shape
    bad = 0
    max_rel = 0.0
    for q in range(n):
        dr, ir, dd, ii = D_ref[q], I_ref[q], D[q], I[q]
        filled_ref = ir >= 0
        filled = ii >= 0
        nfr, nf = int(filled_ref.sum()), int(filled.sum())
        if nfr != nf:
            bad += abs(nfr - nf)
        m = min(nfr, nf)
        if m == 0:
            continue
        scale = np.maximum(np.abs(dr[:m]), atol / rtol)
        rel = np.abs(dr[:m] - dd[:m]) / scale
        max_rel = max(max_rel, float(rel.max()))
        for j in range(m):
            if ir[j] == ii[j]:
                continue
            tol = rtol * max(abs(float(dr[j])), abs(float(dd[j]))) + atol
            where = np.nonzero(ir[:nfr] == ii[j])[0]
            if where.size and abs(float(dr[where[0]]) - float(dd[j])) <= tol and abs(float(dr[where[0]]) - float(dr[j])) <= 2 * tol:
                continue  # near-tie swap inside the list
            if where.size == 0 and abs(float(dd[j]) - float(dr[nfr - 1])) <= 2 * tol and nfr == k:
                continue  # tie at the k boundary
            bad += 1
    return dict(n_id_mismatch_unexplained=bad, max_rel_err=max_rel, n_slots=int(n * k))


def assert_topk_parity(D_ref, I_ref, D, I, rtol=1e-4, atol=1e-6, max_bad_frac=0.0):
    r = compare_topk(D_ref, I_ref, D, I, rtol, atol)
    assert r["max_rel_err"] <= rtol * 1.0001 + 1e-12, r
    assert r["n_id_mismatch_unexplained"] <= max_bad_frac * r["n_slots"], r
    return r


def assert_rerank_parity(f, ix, xq, k, nprobe, R, metric, D, I, filters=None, keys=None, coarse_dis=None):
    """Parity of a has_rank=True search with NO unexplained slack.  Re-ranked distances are bit-identical where ids
    agree.  An id may differ from the reference only for one reason: the two engines nominate slightly different
    recall sets, because their ADC sums round differently (1e-4 relative contract) and the recall set is
    cut at its recall_num-th ADC distance.  Every differing id must therefore sit within the ADC tolerance of that
    cut: the reference's extra id at the boundary of the REFERENCE's recall set (has_rank=False, k=recall_num), ours at
    the boundary of OUR recall set.  Returns the fraction of identical id slots."""
    pj = json.dumps({"nprobe": nprobe, "recall_num": R, "metric_type": metric})
    kw = dict(filters=filters or [])
    if keys is not None:
        kw.update(keys=keys, coarse_dis=coarse_dis)
    D_ref, I_ref = f.ref.search(xq, k, pj, has_rank=True, **kw)
    same = I_ref == I
    assert np.array_equal(D_ref[same], D[same]), "re-ranked distances must be bit-identical where ids agree"
    diff_q = np.nonzero((~same).any(axis=1))[0]
    if diff_q.size:
        Ra = max(R, k)
        Da_ref, Ia_ref = f.ref.search(xq[diff_q], Ra, pj, has_rank=False, **kw)
        rc, Da, Ia = ix.Search(xq[diff_q], Ra, nprobe=nprobe, recall_num=R, metric=metric, has_rank=False, **kw)
        assert rc == 0
        for t, q in enumerate(diff_q):
            # exact-distance (near-)ties at the k boundary or inside the list are ordering freedom, not a difference
            if compare_topk(D_ref[q:q + 1], I_ref[q:q + 1], D[q:q + 1], I[q:q + 1], rtol=1e-6, atol=0.0)[
                    "n_id_mismatch_unexplained"] == 0:
                continue
            for (ids_mine, ids_other, Dset, Iset, who) in ((I_ref[q], I[q], Da_ref[t], Ia_ref[t], "reference"),
                                                           (I[q], I_ref[q], Da[t], Ia[t], "ours")):
                filled = Iset >= 0
                assert filled.all(), "a recall set that is not full has no boundary: ids must be identical (q=%d)" % q
                cut = float(Dset[filled][-1])
                for v in set(ids_mine[ids_mine >= 0].tolist()) - set(ids_other.tolist()):
                    pos = np.nonzero(Iset == v)[0]
                    assert pos.size == 1, "id %d of %s result is not in its own recall set (q=%d)" % (v, who, q)
                    tol = 1e-4 * max(abs(cut), abs(float(Dset[pos[0]]))) + 1e-6
                    assert abs(float(Dset[pos[0]]) - cut) <= tol, (
                        "q=%d id=%d differs but is not at the recall-set boundary of %s (adc %.7g, cut %.7g)" % (
                            q, v, who, float(Dset[pos[0]]), cut))
    return float(same.mean())


# ----------------------------------------------------------------------------------------------
# small reference-built indexes shared by the GPU parity tests
# ----------------------------------------------------------------------------------------------
class RefFixture:
    """A reference (oracle/_ref) IVFPQ index over seeded synthetic data + everything needed to
    mirror it into the device library."""

    def __init__(self, N, d, nlist, M, metric, nq=64, n_clusters=64, seed_shift=0, extra_params=None):
        from gamma_b200 import synth
        from oracle import ref
        self.N, self.d, self.nlist, self.M, self.metric = N, d, nlist, M, metric
        normalize = metric != "L2"
        self.xb = synth.mixture(N, d, synth.SEED_BASE + seed_shift, n_clusters=n_clusters, normalize=normalize)
        self.xq = synth.mixture(nq, d, synth.SEED_QUERY + seed_shift, n_clusters=n_clusters, normalize=normalize)
        mp = {"ncentroids": nlist, "nsubvector": M, "metric_type": metric, "nprobe": 8}
        mp.update(extra_params or {})
        self.model_json = json.dumps(mp)
        self.ref = ref.RefIndex(d, "IVFPQ", self.model_json, indexing_size=N, bitmap_bits=max(N * 2, 1024))
        self.ref.add_raw(self.xb)
        self.ref.indexing()
        self.ref.add_to_index()
        self.centroids = self.ref.centroids()
        self.pq = self.ref.pq_centroids()
        self.opq = self.ref.opq()  # (A, b) when the model was created with "opq"
        self.lists = self.ref.lists()
        self.deleted = []  # docs deleted in the (cached, shared) reference index; mirror() replays them

    def delete(self, docs):
        for doc in docs:
            if int(doc) not in self.deleted:
                self.ref.delete(int(doc))
                self.deleted.append(int(doc))

    def mirror(self, device=0, raw=True):
        """Build the device index from the reference's trained state and postings (list order kept)."""
        from gamma_b200 import api
        ix = api.B200IVFPQ(device)
        assert ix.Init(self.model_json, self.d) == 0, api.lib().gb200_last_error()
        ix.set_quantizers(self.centroids, self.pq)
        if self.opq is not None:
            ix.set_opq(*self.opq)
        list_no = np.concatenate([np.full(len(ids), l, np.int32) for l, (ids, _) in enumerate(self.lists)])
        vids = np.concatenate([ids for ids, _ in self.lists])
        codes = np.concatenate([c for _, c in self.lists])
        # feed in vid order like the engine's AddRTVecsToIndex loop would (per-list order is preserved)
        order = np.argsort(vids, kind="stable")
        assert ix.append(list_no[order], vids[order], codes[order]) == 0, api.lib().gb200_last_error()
        if raw:
            ix.upload_raw(self.xb)
        if self.deleted:
            ix.set_deleted(np.array(self.deleted, np.int64), True)
        return ix


_FIXTURE_CACHE = {}


def get_ref_fixture(key, **kw):
    if key not in _FIXTURE_CACHE:
        _FIXTURE_CACHE[key] = RefFixture(**kw)
    return _FIXTURE_CACHE[key]
