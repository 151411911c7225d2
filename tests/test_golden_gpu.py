"""GPU: the CUDA path against the committed golden fixtures (outputs of the compiled reference)."""
import numpy as np
import pytest

from conftest import assert_topk_parity
from golden_util import GOLDEN, Golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GOLDEN)
def test_cuda_path_matches_golden(name):
    g = Golden(name)
    from gamma_b200 import api
    import json
    ix = api.B200IVFPQ(0)
    mj = json.dumps({"ncentroids": g.nlist, "nsubvector": g.M, "metric_type": g.metric, "nprobe": g.nprobe})
    assert ix.Init(mj, g.d) == 0
    ix.set_quantizers(g["centroids"], g["pq"])
    # Rebuild the realtime lists through the same operations the reference saw: appends in vid order,
    # then the Update that moved one posting (old copy gets the dead flag, new copy appended).
    flat = [(int(i & 0x7fffffffffffffff), l, c, bool(i < 0)) for l, (ids, codes) in enumerate(g.lists)
            for i, c in zip(ids, codes)]
    dead = [t for t in flat if t[3]]
    assert len(dead) == 1
    moved_vid, old_list, old_code, _ = dead[0]
    new = [t for t in flat if t[0] == moved_vid and not t[3]][0]
    first = sorted([t for t in flat if not (t[0] == moved_vid and not t[3])], key=lambda t: t[0])
    assert ix.append(np.array([t[1] for t in first], np.int32), np.array([t[0] for t in first], np.int64),
                     np.stack([t[2] for t in first])) == 0
    assert ix.update(moved_vid, new[1], new[2]) == 0
    for l in range(g.nlist):
        ids, codes = ix.get_list(l)
        assert np.array_equal(ids, g.lists[l][0]) and np.array_equal(codes, g.lists[l][1])
    ix.upload_raw(g["xb"])
    pre = dict(keys=g["coarse_keys"], coarse_dis=g["coarse_dis"])
    cd, keys = ix.coarse(g["xq"], g.nprobe)
    assert (keys == g["coarse_keys"]).mean() > 0.98
    rc, D, I = ix.Search(g["xq"], g.R, recall_num=g.R, metric=g.metric, has_rank=False, **pre)
    assert rc == 0
    assert_topk_parity(g["norank_D"], g["norank_I"], D, I, rtol=1e-4, atol=1e-5)
    rc, D, I = ix.Search(g["xq"], 10, recall_num=g.R, metric=g.metric, has_rank=True, **pre)
    assert_topk_parity(g["rank_D"], g["rank_I"], D, I, rtol=1e-6, atol=0)
    ix.set_deleted(g.deleted)
    rc, D, I = ix.Search(g["xq"], 10, recall_num=g.R, metric=g.metric, has_rank=True, filters=g.filters, **pre)
    assert_topk_parity(g["filt_rank_D"], g["filt_rank_I"], D, I, rtol=1e-6, atol=0)
    rc, D, I = ix.Search(g["xq"], g.R, recall_num=g.R, metric=g.metric, has_rank=False, filters=g.filters[:1], **pre)
    assert_topk_parity(g["filt_norank_D"], g["filt_norank_I"], D, I, rtol=1e-4, atol=1e-5)
    rc, D, I = ix.Search(g["xq"], 10, recall_num=g.R, metric=g.metric, has_rank=True, filters=g.filters,
                         min_score=g.window[0], max_score=g.window[1], **pre)
    assert_topk_parity(g["win_rank_D"], g["win_rank_I"], D, I, rtol=1e-6, atol=0)
    rc, D, I = ix.flat_search(g["xq"], 10, metric=g.metric, filters=g.filters)
    assert rc == 0
    assert np.array_equal(I, g["flat_I"]) and np.array_equal(D, g["flat_D"])  # flat: bit-exact
