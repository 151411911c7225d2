"""GPU unit test of the streaming top-R selection primitive (BlockTopR: append + radix-select prune)
against a host sort — the k-select inside the scan, coarse and flat kernels."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,R,cap,batch,threads", [
    (5000, 100, 1024, 512, 256),     # the headline shape: 4 keys per thread
    (20000, 10, 512, 128, 256),      # flat k=10
    (3000, 600, 2048, 512, 256),     # recall_num 600: 16 keys per thread path
    (900, 600, 2048, 512, 256),      # fewer than cap, more than R: only the final prune
    (500, 600, 2048, 512, 256),      # fewer than R: nothing dropped
    (40000, 2048, 4096, 1024, 256),  # maximum recall_num / nprobe
    (7000, 100, 1024, 256, 384),
])
def test_select_matches_sort(n, R, cap, batch, threads):
    from gamma_b200 import api
    rng = np.random.default_rng(n + R)
    # realistic keys: float distances in a narrow range as the high word, unique scan order as the low word
    dist = rng.normal(30.0, 4.0, n).astype(np.float32).view(np.uint32).astype(np.uint64) | np.uint64(0x80000000)
    keys = (dist << np.uint64(32)) | np.arange(n, dtype=np.uint64)
    got = api.debug_select(keys, R, cap, batch, threads)
    want = np.sort(keys)[:R]
    assert np.array_equal(np.sort(got), want)


def test_select_with_massive_ties_uses_scan_order():
    from gamma_b200 import api
    n, R = 6000, 100
    hi = np.full(n, 0xC1F00000, np.uint64)  # every distance word identical
    hi[::7] = 0xC1E00000                     # ~857 strictly better ones: the R-th falls inside this tie group
    keys = (hi << np.uint64(32)) | np.random.default_rng(1).permutation(n).astype(np.uint64)
    got = api.debug_select(keys, R, 1024, 512, 256)
    assert np.array_equal(np.sort(got), np.sort(keys)[:R])
