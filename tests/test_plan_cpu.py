"""CPU: the positional work plan of the IVFPQ scan (gb200_debug_plan — pure host arithmetic of the C-ABI library).
DESIGN.md §4: whole waves of resident CTAs stay unsplit, only the last partial wave is split, every query owns
`rows` candidate rows that the re-rank can merge (rows * recall_num <= 8192)."""
import ctypes

import pytest

from gamma_b200 import api


def plan(n, slots, nprobe, R, s_uniform=1):
    out = (ctypes.c_int * 4)()
    L = api.lib()
    L.gb200_debug_plan.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p]
    assert L.gb200_debug_plan(n, slots, nprobe, R, s_uniform, out) == 0
    return tuple(out)


@pytest.mark.parametrize("slots", [444, 296])
def test_plan_invariants(slots):
    for nprobe in (1, 5, 32, 64):
        for R in (10, 100, 512, 2048):
            for n in list(range(1, 40)) + [slots - 1, slots, slots + 1, 1000, 1024, 2 * slots, 2 * slots + 7, 4096]:
                for su in (1, 3, 8):
                    n_full, s_tail, n_items, rows = plan(n, slots, nprobe, R, su)
                    assert 0 <= n_full <= n and n_full % slots == 0
                    assert 1 <= s_tail <= max(1, nprobe) and s_tail * R <= max(8192, R)
                    assert n_items == n_full + (n - n_full) * s_tail
                    assert rows == (1 if n_full == n else s_tail)
                    if n >= slots:
                        # the split tail never needs more than one extra wave of CTA slots
                        assert (n - n_full) * s_tail <= max(slots, n - n_full)
                    else:
                        assert n_full == 0 and s_tail == min(su, nprobe, max(1, 8192 // R) if su * R > 8192 else su)


def test_headline_batch_plan():
    # batch 1024 on 148 SMs x 3 CTAs: two full waves of unsplit queries, the remaining 136 split three ways
    assert plan(1024, 444, 32, 100) == (888, 3, 888 + 136 * 3, 3)
    # 2 CTAs per SM (512-thread shape / M = 64 kernel): three full waves, the remaining 136 split two ways
    assert plan(1024, 296, 32, 100) == (888, 2, 888 + 136 * 2, 2)
    assert plan(888, 444, 32, 100) == (888, 1, 888, 1)


def test_bad_arguments_are_rejected():
    out = (ctypes.c_int * 4)()
    assert api.lib().gb200_debug_plan(0, 444, 32, 100, 1, out) != 0
    assert api.lib().gb200_debug_plan(8, 444, 32, 100, 1, None) != 0
